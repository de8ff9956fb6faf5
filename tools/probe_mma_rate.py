"""SS-mode tcgen05.mma issue/throughput probe: cycles per M128 x N x K16 MMA for several N."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pidnet_b200 import _lib
lib = _lib.load_probe()
dev = torch.device('cuda:0')
for blocks in (1, 148):
    for N in (32, 64, 128, 256):
        for distinct in (0, 1):
            out = torch.zeros(blocks, dtype=torch.int64, device=dev)
            iters = 2048
            for _ in range(2):
                _lib.check(lib.pidnet_probe_mma_rate(None, N, iters, distinct, blocks, C.c_void_p(out.data_ptr())))
            cyc = out.float().mean().item() / (iters * 4)
            ideal = 128 * N / 256
            print(f'blocks={blocks:3d} N={N:3d} distinct_operands={distinct}: {cyc:6.1f} cycles/MMA (tensor floor {ideal:.0f}, '
                  f'smem operand wavefronts {(128 + N) * 32 // 128})')
