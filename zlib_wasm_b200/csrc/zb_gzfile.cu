// zb_gzfile.cu — the gz* file layer of zlib.h (gzlib.c, gzread.c, gzwrite.c, gzclose.c) over the
// engine.  Host code only: file I/O, mode parsing, buffering, error bookkeeping.
//
// Reading (gzread.c:76-234 decodes member after member through a 8 KiB window): the file is
// taken in whole, its members are discovered and inflated in ONE batch on the device
// (zb200_gunzip_host: a gzip file has no index), and reads are served from the result.  A file
// that does not start with the gzip magic is handed through unchanged (gzread.c:gz_look
// "direct" mode), bytes after the last member that are no member are ignored.
// Writing (gzwrite.c:11-141): bytes are collected and pushed through deflate() with the gzip
// wrapper in pieces of 64 MiB, on gzflush() and on gzclose(); mode "a" appends a new member,
// "T" writes the bytes through uncompressed (gzwrite.c:gz_init direct).
// Not offered: "+" modes (gzlib.c:130 rejects them too), seeking backwards while writing.
#include "zb_internal.h"
#include "../../include/zb200_zlib.h"
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <string>
#include <vector>

using namespace zb;

namespace {

constexpr uint32_t kGzMagic = 0x5a42475au;
constexpr size_t kWritePiece = (size_t)64 << 20;

struct GzState {
    // the three public fields of zlib.h:1819-1823 (the gzgetc() macro reads them): `have` stays 0 so
    // that the macro always calls the function
    unsigned have; unsigned char *next; long pos;
    uint32_t magic;
    int fd; bool own_fd;
    bool writing, append, direct, exclusive;
    int level, strategy;
    int err; std::string msg;
    // reading
    bool loaded, eof_seen;
    std::vector<uint8_t> data;           // the decoded file
    size_t rpos;
    int unget;                           // pushed-back byte or -1
    long raw_size;
    // writing
    std::vector<uint8_t> wbuf;
    z_stream strm; bool strm_open;
    long wpos;
};

GzState *gstate(gzFile f) {
    GzState *s = reinterpret_cast<GzState *>(f);
    return s && s->magic == kGzMagic ? s : nullptr;
}

void gz_set_error(GzState *s, int err, const char *msg) { s->err = err; s->msg = msg ? msg : ""; }

bool write_all(int fd, const uint8_t *p, size_t n) {
    while (n) {
        const ssize_t k = write(fd, p, n > ((size_t)1 << 30) ? ((size_t)1 << 30) : n);
        if (k < 0) { if (errno == EINTR) continue; return false; }
        p += k; n -= (size_t)k;
    }
    return true;
}

// push wbuf through deflate() (or straight to the file in direct mode)
int gz_drain(GzState *s, int flush) {
    if (s->direct) {
        if (!s->wbuf.empty() && !write_all(s->fd, s->wbuf.data(), s->wbuf.size())) { gz_set_error(s, Z_ERRNO, strerror(errno)); return -1; }
        s->wbuf.clear();
        return 0;
    }
    if (!s->strm_open) {
        memset(&s->strm, 0, sizeof s->strm);
        const int r = deflateInit2_(&s->strm, s->level, Z_DEFLATED, 15 + 16, 8, s->strategy, ZLIB_VERSION, (int)sizeof(z_stream));
        if (r != Z_OK) { gz_set_error(s, r == Z_MEM_ERROR ? Z_MEM_ERROR : Z_STREAM_ERROR, s->strm.msg ? s->strm.msg : "deflateInit2 failed"); return -1; }
        s->strm_open = true;
    }
    std::vector<uint8_t> out((size_t)4 << 20);
    size_t fed = 0;
    for (;;) {
        const size_t left = s->wbuf.size() - fed;
        const uInt step = (uInt)(left > ((size_t)1 << 30) ? ((size_t)1 << 30) : left);
        s->strm.next_in = s->wbuf.data() + fed; s->strm.avail_in = step;
        fed += step;
        const int fl = fed == s->wbuf.size() ? flush : Z_NO_FLUSH;
        int r;
        do {
            s->strm.next_out = out.data(); s->strm.avail_out = (uInt)out.size();
            r = deflate(&s->strm, fl);
            if (r != Z_OK && r != Z_STREAM_END && r != Z_BUF_ERROR) { gz_set_error(s, r, s->strm.msg ? s->strm.msg : "deflate failed"); return -1; }
            const size_t got = out.size() - s->strm.avail_out;
            if (got && !write_all(s->fd, out.data(), got)) { gz_set_error(s, Z_ERRNO, strerror(errno)); return -1; }
        } while (s->strm.avail_out == 0);
        if (fed == s->wbuf.size()) break;
    }
    s->wbuf.clear();
    if (flush == Z_FINISH) { deflateEnd(&s->strm); s->strm_open = false; }
    return 0;
}

// take the whole file in and decode it
int gz_load(GzState *s) {
    if (s->loaded) return 0;
    s->loaded = true;
    std::vector<uint8_t> raw;
    uint8_t tmp[1 << 16];
    for (;;) {
        const ssize_t k = read(s->fd, tmp, sizeof tmp);
        if (k < 0) { if (errno == EINTR) continue; gz_set_error(s, Z_ERRNO, strerror(errno)); return -1; }
        if (k == 0) break;
        raw.insert(raw.end(), tmp, tmp + k);
    }
    s->raw_size = (long)raw.size();
    if (raw.size() < 2 || raw[0] != 0x1f || raw[1] != 0x8b) {   // gzread.c:gz_look: not gzip -> copy through
        s->direct = true;
        s->data.swap(raw);
        return 0;
    }
    zb200_ctx *ctx = zlib_api_ctx();
    if (!ctx) { gz_set_error(s, Z_STREAM_ERROR, "zlib-b200: no usable CUDA device (no CPU path)"); return -1; }
    size_t cap = raw.size() * 4 + 65536, len = 0;
    int st = 0;
    // a file that is ONE member with flush points inside (what gzwrite above produces, pigz -i): its runs in parallel
    if (raw.size() >= ((size_t)1 << 20)) {
        zb200_member_result q;
        for (int attempt = 0; attempt < 2; ++attempt) {
            s->data.resize(cap);
            if (zb200_inflate_stream_host(ctx, raw.data(), raw.size(), ZB200_WRAP_GZIP, s->data.data(), cap, &q) != ZB200_OK) break;
            if (q.status == ZB200_INF_OUTPUT_FULL && q.out_len > cap && attempt == 0) { cap = (size_t)q.out_len; continue; }
            if (q.status == ZB200_INF_OK && q.in_used == raw.size()) { s->data.resize((size_t)q.out_len); return 0; }
            break;
        }
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        s->data.resize(cap);
        const int r = zb200_gunzip_host(ctx, raw.data(), raw.size(), s->data.data(), cap, &len, &st, nullptr, 0, nullptr);
        if (r == ZB200_ERR_OUTPUT && attempt == 0) { cap = len; continue; }
        if (r != ZB200_OK) { s->data.clear(); gz_set_error(s, r == ZB200_ERR_NOMEM ? Z_MEM_ERROR : Z_STREAM_ERROR, zb200_last_error()); return -1; }
        break;
    }
    s->data.resize(len);
    if (st == ZB200_INF_TRUNCATED) gz_set_error(s, Z_BUF_ERROR, "unexpected end of file");      // gzread.c:gz_decomp
    else if (st != ZB200_INF_OK) gz_set_error(s, Z_DATA_ERROR, zb200_inflate_msg(st));
    return 0;
}

gzFile gz_open_common(const char *path, int fd, const char *mode) {
    if (!mode) return nullptr;
    GzState *s = new (std::nothrow) GzState();
    if (!s) return nullptr;
    s->have = 0; s->next = nullptr; s->pos = 0; s->magic = kGzMagic; s->fd = -1; s->own_fd = true;
    s->writing = s->append = s->direct = s->exclusive = false; s->level = Z_DEFAULT_COMPRESSION; s->strategy = Z_DEFAULT_STRATEGY;
    s->err = Z_OK; s->loaded = s->eof_seen = false; s->rpos = 0; s->unget = -1; s->raw_size = 0; s->strm_open = false; s->wpos = 0;
    bool have_mode = false, cloexec = false;
    for (const char *m = mode; *m; ++m) {                      // gzlib.c:113-165
        if (*m >= '0' && *m <= '9') s->level = *m - '0';
        else switch (*m) {
            case 'r': s->writing = false; have_mode = true; break;
            case 'w': s->writing = true; have_mode = true; break;
            case 'a': s->writing = true; s->append = true; have_mode = true; break;
            case '+': delete s; return nullptr;
            case 'b': break;
            case 'e': cloexec = true; break;
            case 'x': s->exclusive = true; break;
            case 'f': s->strategy = Z_FILTERED; break;
            case 'h': s->strategy = Z_HUFFMAN_ONLY; break;
            case 'R': s->strategy = Z_RLE; break;
            case 'F': s->strategy = Z_FIXED; break;
            case 'T': s->direct = true; break;
            default: break;
        }
    }
    if (!have_mode || (!s->writing && s->direct)) { delete s; return nullptr; }   // gzlib.c:168-180: "T" only for writing
    if (path) {
        int fl = s->writing ? (O_WRONLY | O_CREAT | (s->exclusive ? O_EXCL : 0) | (s->append ? O_APPEND : O_TRUNC)) : O_RDONLY;
        if (cloexec) fl |= O_CLOEXEC;
        s->fd = open(path, fl, 0666);
        if (s->fd < 0) { delete s; return nullptr; }
    } else {
        s->fd = fd;
        if (fd < 0) { delete s; return nullptr; }
    }
    return reinterpret_cast<gzFile>(s);
}

}  // namespace

extern "C" {

gzFile gzopen(const char *path, const char *mode) { return path ? gz_open_common(path, -1, mode) : nullptr; }
gzFile gzdopen(int fd, const char *mode) { return gz_open_common(nullptr, fd, mode); }
int gzbuffer(gzFile file, unsigned size) { GzState *s = gstate(file); (void)size; return s && !s->loaded && s->wbuf.empty() && !s->wpos ? 0 : -1; }

int gzsetparams(gzFile file, int level, int strategy) {      // gzwrite.c:gzsetparams
    GzState *s = gstate(file);
    if (!s || !s->writing || s->err != Z_OK || s->direct) return Z_STREAM_ERROR;
    if (level == s->level && strategy == s->strategy) return Z_OK;
    if (s->strm_open) {
        if (gz_drain(s, Z_BLOCK) < 0) return s->err;
        const int r = deflateParams(&s->strm, level, strategy);
        if (r != Z_OK) return r;
    }
    s->level = level; s->strategy = strategy;
    return Z_OK;
}

int gzread(gzFile file, voidp buf, unsigned len) {
    GzState *s = gstate(file);
    if (!s || s->writing) return -1;
    if ((int)len < 0) { gz_set_error(s, Z_STREAM_ERROR, "request does not fit in an int"); return -1; }   // gzread.c:gzread
    if (gz_load(s) < 0) return -1;
    uint8_t *o = (uint8_t *)buf;
    unsigned got = 0;
    if (len && s->unget >= 0) { o[got++] = (uint8_t)s->unget; s->unget = -1; }
    const size_t left = s->data.size() - s->rpos;
    const size_t k = left < len - got ? left : len - got;
    if (k) memcpy(o + got, s->data.data() + s->rpos, k);
    s->rpos += k; got += (unsigned)k;
    s->pos += got;
    if (got < len) s->eof_seen = true;
    if (got == 0 && len && s->err != Z_OK && s->err != Z_BUF_ERROR) return -1;        // data error reached: gzread.c returns -1
    return (int)got;
}

z_size_t gzfread(voidp buf, z_size_t size, z_size_t nitems, gzFile file) {
    GzState *s = gstate(file);
    if (!s || s->writing || !size) return 0;
    z_size_t total = size * nitems, done = 0;
    if (nitems && total / nitems != size) { gz_set_error(s, Z_STREAM_ERROR, "request does not fit in a size_t"); return 0; }
    while (done < total) {
        const unsigned step = (unsigned)(total - done > 0x40000000u ? 0x40000000u : total - done);
        const int k = gzread(file, (uint8_t *)buf + done, step);
        if (k <= 0) break;
        done += (z_size_t)k;
        if ((unsigned)k < step) break;
    }
    return done / size;
}

int gzwrite(gzFile file, voidpc buf, unsigned len) {
    GzState *s = gstate(file);
    if (!s || !s->writing || s->err != Z_OK) return 0;
    if ((int)len < 0) { gz_set_error(s, Z_DATA_ERROR, "requested length does not fit in int"); return 0; }   // gzwrite.c:gzwrite
    if (len == 0) return 0;
    s->wbuf.insert(s->wbuf.end(), (const uint8_t *)buf, (const uint8_t *)buf + len);
    s->wpos += len; s->pos = s->wpos;
    if (s->wbuf.size() >= kWritePiece && gz_drain(s, Z_NO_FLUSH) < 0) return 0;
    return (int)len;
}

z_size_t gzfwrite(voidpc buf, z_size_t size, z_size_t nitems, gzFile file) {
    GzState *s = gstate(file);
    if (!s || !s->writing || !size) return 0;
    z_size_t total = size * nitems, done = 0;
    if (nitems && total / nitems != size) { gz_set_error(s, Z_STREAM_ERROR, "request does not fit in a size_t"); return 0; }
    while (done < total) {
        const unsigned step = (unsigned)(total - done > 0x40000000u ? 0x40000000u : total - done);
        if (gzwrite(file, (const uint8_t *)buf + done, step) != (int)step) break;
        done += step;
    }
    return done / size;
}

int gzputs(gzFile file, const char *str) {
    if (!str) return -1;
    const size_t n = strlen(str);
    if ((int)n < 0) return -1;
    const int k = gzwrite(file, str, (unsigned)n);
    return k == 0 && n != 0 ? -1 : k;
}

int gzputc(gzFile file, int c) { unsigned char b = (unsigned char)c; return gzwrite(file, &b, 1) == 1 ? (int)b : -1; }

int gzvprintf(gzFile file, const char *format, va_list va) {   // gzwrite.c:gzvprintf
    GzState *s = gstate(file);
    if (!s || !s->writing || !format) return Z_STREAM_ERROR;
    va_list vb;
    va_copy(vb, va);
    const int n = vsnprintf(nullptr, 0, format, va);
    if (n <= 0) { va_end(vb); return n; }
    std::vector<char> tmp((size_t)n + 1);
    vsnprintf(tmp.data(), tmp.size(), format, vb);
    va_end(vb);
    return gzwrite(file, tmp.data(), (unsigned)n);
}

int gzprintf(gzFile file, const char *format, ...) {
    va_list va;
    va_start(va, format);
    const int n = gzvprintf(file, format, va);
    va_end(va);
    return n;
}

char *gzgets(gzFile file, char *buf, int len) {              // gzread.c:gzgets
    GzState *s = gstate(file);
    if (!s || s->writing || !buf || len < 1) return nullptr;
    if (gz_load(s) < 0) return nullptr;
    int n = 0;
    while (n < len - 1) {
        int c;
        if (s->unget >= 0) { c = s->unget; s->unget = -1; }
        else if (s->rpos < s->data.size()) c = s->data[s->rpos++];
        else { s->eof_seen = true; break; }
        buf[n++] = (char)c; s->pos++;
        if (c == '\n') break;
    }
    if (n == 0) return nullptr;
    buf[n] = 0;
    return buf;
}

int (gzgetc)(gzFile file) { unsigned char b; return gzread(file, &b, 1) == 1 ? (int)b : -1; }
int gzgetc_(gzFile file) { return (gzgetc)(file); }

int gzungetc(int c, gzFile file) {
    GzState *s = gstate(file);
    if (!s || s->writing || c < 0 || s->unget >= 0) return -1;
    if (gz_load(s) < 0) return -1;
    s->unget = c & 0xff; s->pos--; s->eof_seen = false;
    return c;
}

int gzflush(gzFile file, int flush) {
    GzState *s = gstate(file);
    if (!s || !s->writing || s->err != Z_OK) return Z_STREAM_ERROR;
    if (flush < 0 || flush > Z_FINISH) return Z_STREAM_ERROR;
    return gz_drain(s, flush) < 0 ? s->err : Z_OK;
}

z_off_t gzseek(gzFile file, z_off_t offset, int whence) {   // gzlib.c:gzseek64
    GzState *s = gstate(file);
    if (!s || (whence != SEEK_SET && whence != SEEK_CUR)) return -1;
    if (s->writing) {
        const long target = whence == SEEK_SET ? offset : s->wpos + offset;
        if (target < s->wpos) return -1;
        std::vector<uint8_t> zeros((size_t)(target - s->wpos), 0);
        if (!zeros.empty() && gzwrite(file, zeros.data(), (unsigned)zeros.size()) != (int)zeros.size()) return -1;
        return s->wpos;
    }
    if (gz_load(s) < 0) return -1;
    const long cur = (long)s->rpos - (s->unget >= 0 ? 1 : 0);
    const long target = whence == SEEK_SET ? offset : cur + offset;
    if (target < 0) return -1;
    s->unget = -1;
    s->rpos = (size_t)target > s->data.size() ? s->data.size() : (size_t)target;
    s->pos = (long)s->rpos; s->eof_seen = false;
    if (s->err == Z_BUF_ERROR) gz_set_error(s, Z_OK, nullptr);
    return s->pos;
}

int gzrewind(gzFile file) {
    GzState *s = gstate(file);
    if (!s || s->writing) return -1;
    return gzseek(file, 0, SEEK_SET) == 0 ? 0 : -1;
}

z_off_t gztell(gzFile file) { GzState *s = gstate(file); return s ? s->pos : -1; }

z_off_t gzoffset(gzFile file) {
    GzState *s = gstate(file);
    if (!s) return -1;
    const off_t o = lseek(s->fd, 0, SEEK_CUR);
    return o < 0 ? -1 : (z_off_t)o;
}

// the LFS names zlib.h maps gzopen / gzseek / gztell / gzoffset to under _FILE_OFFSET_BITS=64 (zlib.h:1893-1912, gzlib.c:268,342):
// z_off_t is 64 bits wide on this platform, so they are the same functions
gzFile gzopen64(const char *path, const char *mode) { return gzopen(path, mode); }
long gzseek64(gzFile file, long offset, int whence) { return gzseek(file, offset, whence); }
long gztell64(gzFile file) { return gztell(file); }
long gzoffset64(gzFile file) { return gzoffset(file); }

int gzeof(gzFile file) { GzState *s = gstate(file); return s && !s->writing && s->eof_seen ? 1 : 0; }
int gzdirect(gzFile file) { GzState *s = gstate(file); if (!s) return 0; if (!s->writing) gz_load(s); return s->direct ? 1 : 0; }

int gzclose(gzFile file) {
    GzState *s = gstate(file);
    if (!s) return Z_STREAM_ERROR;
    int ret = Z_OK;
    if (s->writing) {
        if (s->err == Z_OK) { if (gz_drain(s, Z_FINISH) < 0) ret = s->err; }
        else ret = s->err;
        if (s->strm_open) deflateEnd(&s->strm);
    } else if (s->err == Z_BUF_ERROR) ret = Z_BUF_ERROR;       // gzclose.c / gzread.c:gzclose_r
    if (s->own_fd && close(s->fd) != 0 && ret == Z_OK) ret = Z_ERRNO;
    s->magic = 0;
    delete s;
    return ret;
}
int gzclose_r(gzFile file) { GzState *s = gstate(file); return s && !s->writing ? gzclose(file) : Z_STREAM_ERROR; }
int gzclose_w(gzFile file) { GzState *s = gstate(file); return s && s->writing ? gzclose(file) : Z_STREAM_ERROR; }

const char *gzerror(gzFile file, int *errnum) {
    GzState *s = gstate(file);
    if (!s) return nullptr;
    if (errnum) *errnum = s->err;
    return s->err == Z_MEM_ERROR ? "out of memory" : s->msg.c_str();
}

void gzclearerr(gzFile file) {
    GzState *s = gstate(file);
    if (!s) return;
    if (!s->writing) s->eof_seen = false;
    gz_set_error(s, Z_OK, nullptr);
}

}  // extern "C"
