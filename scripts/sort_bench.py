"""Stand-alone timing of the onesweep sort at 16M keys (scripts/sort_bench.py [n] [bits])."""
import sys, numpy as np, ctypes as C
sys.path.insert(0, ".")
import torch
from realtimeparticles_b200 import _abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 22
h = _abi.Handle(_abi.BOIDS, 1024, 0)
rng = np.random.default_rng(0)
# cell-id-like keys: particles of a dam break are nearly sorted by x already; use random keys (worst case for the scatter)
keys = torch.from_numpy(rng.integers(0, 3456000, size=n, dtype=np.int64).astype(np.int32)).cuda()
ko, po = torch.empty_like(keys), torch.empty_like(keys)
L = h.L
L.rtp_sort_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
stream = torch.cuda.ExternalStream(h.stream())
for _ in range(3):
    assert L.rtp_sort_keys(h.h, keys.data_ptr(), ko.data_ptr(), po.data_ptr(), n, bits) == 0
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
for a, b in ev:
    a.record(stream)
    L.rtp_sort_keys(h.h, keys.data_ptr(), ko.data_ptr(), po.data_ptr(), n, bits)
    b.record(stream)
torch.cuda.synchronize()
ms = min(a.elapsed_time(b) for a, b in ev)
passes = (bits + 7) // 8
print("n=%d bits=%d passes=%d: %.3f ms (incl. cudaMalloc/copy of the stand-alone entry point); algorithmic %.0f MB -> %.0f GB/s" % (
    n, bits, passes, ms, (4 + 16 * passes) * n / 1e6, (4 + 16 * passes) * n / ms / 1e6))
assert torch.equal(ko.cpu(), torch.sort(keys.cpu(), stable=True).values)
