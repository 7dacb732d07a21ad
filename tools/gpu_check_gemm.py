"""GPU bring-up check for the tcgen05 GEMM: every operand-major combination, ragged sizes,
bias/relu/accumulate, both output types, compared with torch.matmul on the same bf16 inputs."""
import importlib.util
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("las_lib", os.path.join(ROOT, "semi-supervised-asr_b200", "_lib.py"))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)


def run(M, N, K, a_mn, b_mn, c_bf16, bias, relu, acc):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev, generator=g).to(torch.bfloat16)
    ref = A.float() @ B.float().t()
    bias_t = torch.randn(N, device=dev, generator=g) if bias else None
    if bias:
        ref = ref + bias_t
    if relu:
        ref = ref.relu()
    C0 = torch.randn(M, N, device=dev, generator=g)
    if acc:
        ref = ref + (C0.to(torch.bfloat16).float() if c_bf16 else C0)
    Kp = (K + 7) // 8 * 8
    Mp = (M + 7) // 8 * 8
    Np = (N + 7) // 8 * 8
    if a_mn:
        Ast = torch.zeros(K, Mp, device=dev, dtype=torch.bfloat16)
        Ast[:, :M] = A.t()
        lda = Mp
    else:
        Ast = torch.zeros(M, Kp, device=dev, dtype=torch.bfloat16)
        Ast[:, :K] = A
        lda = Kp
    if b_mn:
        Bst = torch.zeros(K, Np, device=dev, dtype=torch.bfloat16)
        Bst[:, :N] = B.t()
        ldb = Np
    else:
        Bst = torch.zeros(N, Kp, device=dev, dtype=torch.bfloat16)
        Bst[:, :K] = B
        ldb = Kp
    C = (C0.to(torch.bfloat16) if c_bf16 else C0.clone()).contiguous()
    L.call("las_gemm_bf16", L.ptr(Ast), int(lda), int(a_mn), L.ptr(Bst), int(ldb), int(b_mn),
           L.ptr(C), int(N), int(c_bf16), L.ptr(bias_t), int(M), int(N), int(K), int(relu),
           int(acc))
    torch.cuda.synchronize()
    err = (C.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    tol = 2e-2 if c_bf16 else 2e-3
    ok = err / scale < tol
    print(f"gemm M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bf16out={c_bf16} bias={bias} relu={relu} acc={acc}: "
          f"max_err={err:.3e} rel={err / scale:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def main():
    print("device:", torch.cuda.get_device_name(0), flush=True)
    ok = True
    shapes = [(128, 128, 64), (256, 256, 256), (300, 200, 136), (1000, 34, 640), (4000, 1280, 256),
              (129, 65, 72), (64, 2560, 1280)]
    for (M, N, K) in shapes:
        ok &= run(M, N, K, 0, 0, 0, False, False, False)
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            ok &= run(512, 384, 320, a_mn, b_mn, 0, True, False, False)
            ok &= run(333, 130, 200, a_mn, b_mn, 0, False, False, True)
    ok &= run(777, 320, 1280, 0, 0, 1, True, True, False)
    ok &= run(777, 320, 1280, 0, 0, 1, True, False, True)
    ok &= run(2560, 256, 8000, 1, 1, 0, False, False, False)
    # timing at the layer-0 input projection shape (config 2): [32000,256] x [2560,256]^T -> f32
    M, N, K = 32000, 2560, 256
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda")
    for c_bf16, Ct in ((0, C), (1, C.to(torch.bfloat16))):
        for _ in range(3):
            L.call("las_gemm_bf16", L.ptr(A), int(K), int(0), L.ptr(B), int(K), int(0), L.ptr(Ct),
                   int(N), int(c_bf16), L.ptr(None), int(M), int(N), int(K), int(0), int(0),
                   L.stream_ptr())
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            L.call("las_gemm_bf16", L.ptr(A), int(K), int(0), L.ptr(B), int(K), int(0), L.ptr(Ct),
                   int(N), int(c_bf16), L.ptr(None), int(M), int(N), int(K), int(0), int(0),
                   L.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byts = M * K * 2 + N * K * 2 + M * N * (2 if c_bf16 else 4)
        print(f"inproj gemm bf16out={c_bf16}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s  {byts / ms / 1e6:.0f} GB/s", flush=True)
    t0 = time.time()
    for _ in range(5):
        ref = A.float() @ B.float().t()
    torch.cuda.synchronize()
    # big square for tensor-pipe rate
    M = N = K = 4096
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        L.call("las_gemm_bf16", L.ptr(A), int(K), int(0), L.ptr(B), int(K), int(0), L.ptr(C),
               int(N), int(1), L.ptr(None), int(M), int(N), int(K), int(0), int(0), L.stream_ptr())
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.call("las_gemm_bf16", L.ptr(A), int(K), int(0), L.ptr(B), int(K), int(0), L.ptr(C),
               int(N), int(1), L.ptr(None), int(M), int(N), int(K), int(0), int(0), L.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"4096^3 gemm: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    print("GEMM CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
