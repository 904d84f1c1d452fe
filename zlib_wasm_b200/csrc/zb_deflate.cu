#include "zb_internal.h"
namespace zb { int deflate_init(zb200_ctx *) { return ZB200_OK; } }
extern "C" {
size_t zb200_deflate_bound(size_t n, size_t, int) { return n + (n >> 12) + (n >> 14) + (n >> 25) + 13; }
size_t zb200_deflate_scratch_bytes(size_t, size_t) { return 0; }
int zb200_deflate_dev(zb200_ctx *, const void *, size_t, size_t, int, int, int, int, void *, size_t, uint64_t *, uint64_t *, void *) {
    zb::set_error("deflate: not built yet"); return ZB200_ERR_PARAM; }
int zb200_deflate_host(zb200_ctx *, const void *, size_t, size_t, int, int, int, int, void *, size_t *, uint32_t *, uint32_t *) {
    zb::set_error("deflate: not built yet"); return ZB200_ERR_PARAM; }
}
