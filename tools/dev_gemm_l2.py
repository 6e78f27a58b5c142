"""Dev: DiT GEMM shapes under the L2 knobs of gemm_sm100.cu (VGPA_GEMM_GROUP_M / _STREAM_OUT / _HINTS are read once per
process, so run one process per variant). Prints CUDA-event times; run under `ncu --metrics dram__bytes_read.sum,...` for traffic."""
import os, sys, torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
BF = torch.bfloat16
M = 35552
shapes = {"qkv": (9216, 3072, dense.EPI_BIAS), "ff1": (12288, 3072, dense.EPI_BIAS_GELU), "ff2": (3072, 12288, dense.EPI_BIAS)}
which = sys.argv[1:] or list(shapes)
tag = f"G={os.environ.get('VGPA_GEMM_GROUP_M', 'dflt')} S={os.environ.get('VGPA_GEMM_STREAM_OUT', '0')} H={os.environ.get('VGPA_GEMM_HINTS', '0')}"
for name in which:
    N, K, epi = shapes[name]
    a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") * 0.02).to(BF); b = torch.zeros(N, device="cuda", dtype=BF)
    out = torch.empty(M, N, device="cuda", dtype=BF)
    for _ in range(2):
        dense.linear(a, w, b, out=out, epilogue=epi)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dense.linear(a, w, b, out=out, epilogue=epi)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{tag} {name} M={M} N={N} K={K}: {ms:.3f} ms {2.0 * M * N * K / ms / 1e9:.0f} TF/s", flush=True)
    del a, w, b, out
