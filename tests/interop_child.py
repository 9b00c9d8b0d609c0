"""Child process of tests/test_interop_gpu.py: plays the CONSUMER of the exported cull outputs (what the renderer's draw does through
VkImportMemoryFdInfoKHR): maps the two exported allocations from their file descriptors, orders itself behind the producer's
interprocess fence, reads the count and the records IN PLACE (no copy made by the producer) and prints count + sha256.

    python tests/interop_child.py <device> <draws_fd> <draws_alloc_bytes> <counts_fd> <counts_alloc_bytes> <count_offset> <rec_bytes> <fence_hex>
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blitzen_b200 import capi  # noqa: E402


def main():
    dev, dfd, dbytes, cfd, cbytes, coff, rec_bytes = (int(x) for x in sys.argv[1:8])
    fence = bytes.fromhex(sys.argv[8])
    lib = capi.load_library()

    def check(rc):
        if rc != 0:
            raise SystemExit("interop child: " + lib.blz_cull_last_error().decode())

    draws, counts = C.c_void_p(), C.c_void_p()
    check(lib.blz_interop_import(dev, dfd, dbytes, C.byref(draws)))
    check(lib.blz_interop_import(dev, cfd, cbytes, C.byref(counts)))
    fbuf = (C.c_ubyte * 64).from_buffer_copy(fence)
    check(lib.blz_interop_wait_fence(fbuf, None))              # default stream waits for the producer's cull passes
    cnt = np.zeros(2, dtype=np.uint32)
    check(lib.blz_interop_read(cnt.ctypes.data_as(C.c_void_p), C.c_void_p(counts.value + coff), 8, None))
    rec = np.zeros(int(cnt[0]) * rec_bytes, dtype=np.uint8)
    if len(rec):
        check(lib.blz_interop_read(rec.ctypes.data_as(C.c_void_p), draws, len(rec), None))
    print("INTEROP", int(cnt[0]), int(cnt[1]), hashlib.sha256(rec.tobytes()).hexdigest(), flush=True)
    check(lib.blz_interop_release(draws, dbytes))
    check(lib.blz_interop_release(counts, cbytes))


if __name__ == "__main__":
    main()
