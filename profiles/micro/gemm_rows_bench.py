#!/usr/bin/env python
"""case_gemm_rows_tc against cuBLAS (torch.mm, bf16 in / fp32 out) on the Linear shapes of the pre-decode producers at the
BASELINE shape (M = 64 x 10 x 256 passage tokens).  usage: python profiles/micro/gemm_rows_bench.py [N K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from case_rg_b200 import _lib as L                       # noqa: E402
from case_rg_b200.producers import _Linear               # noqa: E402


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device('cuda')
    M = 64 * 10 * 256
    st = torch.cuda.current_stream().cuda_stream
    print(f'{"N":>6} {"K":>6} {"epilogue":>22} | {"own ms":>8} {"TF/s":>7} | {"cuBLAS ms":>9} {"TF/s":>7}')
    for N, K, act, res, out32 in ((768, 256, 0, None, False), (256, 256, 0, 'f32', True), (256, 256, 1, None, False),
                                  (3840, 1280, 0, None, False), (1280, 1280, 0, 'bf16', False), (256, 1280, 2, None, False),
                                  (256, 256, 0, None, True)):
        if len(sys.argv) > 2 and (N, K) != (int(sys.argv[1]), int(sys.argv[2])):
            continue
        x = torch.randn(M, K, device=dev).bfloat16()
        lin = _Linear(torch.randn(N, K) / K ** 0.5, torch.randn(N) * 0.1, dev)
        r = None if res is None else (torch.randn(M, N, device=dev) if res == 'f32' else torch.randn(M, N, device=dev).bfloat16())
        y = torch.empty(M, N, dtype=torch.float32 if out32 else torch.bfloat16, device=dev)

        def own():
            L.call('case_gemm_rows_tc', x.data_ptr(), lin.wp.data_ptr(), lin.b.data_ptr(), M, N, K, act, L.ptr(r),
                   (L.BF16 if res == 'bf16' else L.F32) if res else 0, None, y.data_ptr(), L.F32 if out32 else L.BF16, st)
        wt = lin.w16.t().contiguous()
        ms = timeit(own)
        ms2 = timeit(lambda: torch.mm(x, wt, out_dtype=torch.float32))
        fl = 2.0 * M * N * K
        ep = f'act={act} res={res} out={"f32" if out32 else "bf16"}'
        print(f'{N:6d} {K:6d} {ep:>22} | {ms:8.3f} {fl / ms / 1e9:7.0f} | {ms2:9.3f} {fl / ms2 / 1e9:7.0f}')


if __name__ == '__main__' and not (len(sys.argv) > 1 and sys.argv[1] == 'ffn'):
    main()


def ffn_main():
    """The fused feed-forward against its two separate launches (same shapes as the producers' layers)."""
    dev = torch.device('cuda')
    M = 64 * 10 * 256
    st = torch.cuda.current_stream().cuda_stream
    print(f'{"K1":>6} {"epilogue":>22} | {"fused ms":>8} | {"lin1 + lin2 ms":>14}')
    for K1, act, res in ((256, 1, 'f32'), (256, 2, None), (1280, 2, None)):
        x = torch.randn(M, K1, device=dev).bfloat16()
        l1 = _Linear(torch.randn(256, K1) / K1 ** 0.5, torch.randn(256) * 0.1, dev)
        l2 = _Linear(torch.randn(256, 256) / 16, torch.randn(256) * 0.1, dev)
        r = torch.randn(M, 256, device=dev) if res else None
        rm = torch.ones(M, dtype=torch.uint8, device=dev)
        y = torch.empty(M, 256, device=dev)
        h = torch.empty(M, 256, dtype=torch.bfloat16, device=dev)

        def fused():
            L.call('case_ffn_rows_tc', x.data_ptr(), l1.wp.data_ptr(), l1.b.data_ptr(), K1, act, l2.wp.data_ptr(), l2.b.data_ptr(), M,
                   L.ptr(r), L.F32 if res else 0, rm.data_ptr(), y.data_ptr(), L.F32, st)

        def two():
            L.call('case_gemm_rows_tc', x.data_ptr(), l1.wp.data_ptr(), l1.b.data_ptr(), M, 256, K1, act, None, 0, None, h.data_ptr(),
                   L.BF16, st)
            L.call('case_gemm_rows_tc', h.data_ptr(), l2.wp.data_ptr(), l2.b.data_ptr(), M, 256, 256, 0, L.ptr(r), L.F32 if res else 0,
                   rm.data_ptr(), y.data_ptr(), L.F32, st)
        print(f'{K1:6d} {"act=%d res=%s" % (act, res):>22} | {timeit(fused):8.3f} | {timeit(two):14.3f}')


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'ffn':
    ffn_main()
