"""Linear-layer micro-benchmark at the BASELINE shapes: the tcgen05 kernels (rsc_linear_*) against what they replace
(library GEMM + rsc_bias_act_* + rsc_colsum).  CUDA-event medians, L2 flushed between iterations.
   python tools/gbench.py [tag]"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscotr_b200 import _lib, ops  # noqa: E402
from tools.kbench import timeit  # noqa: E402

SHAPES = [  # (name, M tokens, N out, K in, act)
    ('s0.qkv', 640000, 288, 96, 0), ('s0.proj', 640000, 96, 96, 0), ('s0.fc1', 640000, 384, 96, 1), ('s0.fc2', 640000, 96, 384, 0),
    ('s1.qkv', 160000, 576, 192, 0), ('s1.fc1', 160000, 768, 192, 1), ('s1.fc2', 160000, 192, 768, 0),
    ('s2.qkv', 40000, 1152, 384, 0), ('s2.fc1', 40000, 1536, 384, 1), ('s2.fc2', 40000, 384, 1536, 0),
    ('s3.fc1', 10000, 3072, 768, 1), ('s3.fc2', 10000, 768, 3072, 0),
    ('enc.ffn1', 26588, 2048, 256, 2), ('enc.ffn2', 26588, 256, 2048, 0), ('enc.proj', 26588, 256, 256, 0),
]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'gbench'
    dev = 'cuda'
    st = lambda: torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, M, N, K, act in SHAPES:
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
        b = torch.randn(N, device=dev)
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        h = torch.empty_like(y)
        dy = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        dw = torch.zeros(N, K, device=dev)
        db = torch.zeros(N, device=dev)
        code = {0: None, 1: ops.ACT_GELU, 2: ops.ACT_RELU}[act]
        # forward
        t_own = timeit(lambda: _lib.call('rsc_linear_fwd', x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(),
                                         h.data_ptr() if act == 1 else None, M, N, K, K, K, N, act, st()), flush=flush)
        if code is None:
            t_lib = timeit(lambda: F.linear(x, w, b.bfloat16()), flush=flush)
        else:
            t_lib = timeit(lambda: ops.bias_act(F.linear(x, w), b, code), flush=flush)
        fbytes = 2 * (M * K + N * K + M * N * (2 if act == 1 else 1))
        rows.append(dict(k='fwd', name=name, M=M, N=N, K=K, act=act, own_us=round(t_own * 1e3, 1), lib_us=round(t_lib * 1e3, 1),
                         own_gbs=round(fbytes / t_own / 1e6), tflops=round(2 * M * N * K / t_own / 1e9, 1)))
        # dX: the input gradient of THIS layer (dy (M,N) -> dx (M,K)); the activation gradient belongs to the layer BELOW,
        # so it is benchmarked on the transposed role: act on the output of the dX GEMM
        aux = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        for a2 in ([0] if name.endswith('fc2') is False and not name.endswith('ffn2') else [0, 1 if 'fc2' in name else 2]):
            t_own = timeit(lambda: _lib.call('rsc_linear_dx', dy.data_ptr(), w.data_ptr(), aux.data_ptr() if a2 else None,
                                             dx.data_ptr(), M, N, K, N, K, K, a2, st()), flush=flush)
            if a2 == 0:
                t_lib = timeit(lambda: torch.mm(dy, w), flush=flush)
            else:
                hh = aux.clone().requires_grad_(True)
                bb = torch.zeros(K, device=dev, requires_grad=True)
                yy = ops.bias_act(hh, bb, ops.ACT_GELU if a2 == 1 else ops.ACT_RELU)
                g = torch.mm(dy, w)

                def lib_bwd():
                    g2 = torch.mm(dy, w)
                    torch.autograd.grad(yy, (hh, bb), g2, retain_graph=True)
                t_lib = timeit(lib_bwd, flush=flush)
            rows.append(dict(k='dx', name=name, act=a2, own_us=round(t_own * 1e3, 1), lib_us=round(t_lib * 1e3, 1),
                             tflops=round(2 * M * N * K / t_own / 1e9, 1)))
        # dW + db
        t_own = timeit(lambda: _lib.call('rsc_linear_dw', dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), M, N, K, N, K, K,
                                         st()), flush=flush)

        def lib_dw():
            torch.addmm(dw, dy.t(), x, out_dtype=torch.float32, out=dw)
            ops.colsum(dy, out=db)
        t_lib = timeit(lib_dw, flush=flush)
        rows.append(dict(k='dw', name=name, own_us=round(t_own * 1e3, 1), lib_us=round(t_lib * 1e3, 1),
                         own_gbs=round(2 * (M * K + M * N) / t_own / 1e6), tflops=round(2 * M * N * K / t_own / 1e9, 1)))
        del x, y, h, dy, dx, aux
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/%s.jsonl' % tag, 'w') as f:
        for r in rows:
            print(json.dumps(r))
            f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
