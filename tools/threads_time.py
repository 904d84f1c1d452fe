"""Host threads calling the zlib.h names at once: python tools/threads_time.py [KiB per call] [calls per thread]
— aggregate throughput of compress2 (level 6) / uncompress / crc32 calls on pageable buffers for 1, 2, 4 and 8 threads, with
the pool of engine contexts ($ZB200_CONTEXTS, default 4) and with one context (ZB200_CONTEXTS=1: every call waits for the one
before it).  Each configuration runs in its own process (the knob is read once)."""
import ctypes as C
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import refz
    import zlib_wasm_b200 as zb
    kib, calls = int(sys.argv[2]), int(sys.argv[3])
    z = refz.ZlibBinding(zb.LIB_PATH, "")
    n = kib << 10
    warm = refz.gen(n, refz.GEN_TEXT, seed=99)

    def warm_up():
        cap = z.compressBound(n)
        dst, dl, back, bl = C.create_string_buffer(cap), C.c_ulong(cap), C.create_string_buffer(n), C.c_ulong(n)
        z.compress2(dst, C.byref(dl), warm, n, 6); z.uncompress(back, C.byref(bl), dst, dl.value); z.crc32(0, warm, n)

    for _ in range(2):                                           # every slot of the pool: context created, buffers grown
        ths = [threading.Thread(target=warm_up) for _ in range(16)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    for nt in (1, 2, 4, 8):
        data = [refz.gen(n, refz.GEN_TEXT, seed=t + 1) for t in range(nt)]
        res = {}

        def work(t, what):
            d = data[t]
            cap = z.compressBound(n)
            dst, dl = C.create_string_buffer(cap), C.c_ulong(cap)
            z.compress2(dst, C.byref(dl), d, n, 6)
            back = C.create_string_buffer(n)
            for _ in range(calls):
                if what == "compress2":
                    dl.value = cap
                    assert z.compress2(dst, C.byref(dl), d, n, 6) == 0
                elif what == "uncompress":
                    bl = C.c_ulong(n)
                    assert z.uncompress(back, C.byref(bl), dst, dl.value) == 0
                else:
                    z.crc32(0, d, n)

        for what in ("compress2", "uncompress", "crc32"):
            ths = [threading.Thread(target=work, args=(t, what)) for t in range(nt)]
            t0 = time.perf_counter()
            for th in ths:
                th.start()
            for th in ths:
                th.join()
            res[what] = nt * calls * n / (time.perf_counter() - t0) / 1e9
        print("  %d thread(s): compress2 %.2f GB/s, uncompress %.2f GB/s, crc32 %.2f GB/s" % (nt, res["compress2"], res["uncompress"], res["crc32"]), flush=True)
    sys.exit(0)

kib = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for knob in ("1", "4"):
    env = dict(os.environ)
    env["ZB200_CONTEXTS"] = knob
    print("ZB200_CONTEXTS=%s, %d KiB per call, %d calls per thread and kind:" % (knob, kib, calls), flush=True)
    subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(kib), str(calls)], env=env, check=False)
