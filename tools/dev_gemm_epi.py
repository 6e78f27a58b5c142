"""Dev: standalone timing of the GEMM epilogue variants at the DiT shapes (not part of the product)."""
import sys
import torch
sys.path.insert(0, ".")
from videogpa_b200 import dense
BF = torch.bfloat16
M, S, St = 35552, 17776, 226
def t(fn, it=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for (N, K) in [(3072, 3072), (3072, 12288), (9216, 3072), (12288, 3072)]:
    a = torch.randn(M, K, device="cuda").to(BF)
    w = (torch.randn(N, K, device="cuda") * 0.02).to(BF)
    b = torch.zeros(N, device="cuda", dtype=BF)
    out = torch.randn(M, N, device="cuda").to(BF)
    gate = torch.randn(2, 6 * 3072, device="cuda").to(BF)
    fl = 2.0 * M * N * K
    ms0 = t(lambda: dense.linear(a, w, b, out=out, epilogue=dense.EPI_BIAS))
    line = f"N={N} K={K}: bias {ms0:.3f} ms {fl/ms0/1e9:.0f} TF/s"
    if N == 3072:
        ms2 = t(lambda: dense.linear(a, w, b, out=out, epilogue=dense.EPI_GATE_RES, rows_per_sample=S, text_rows=St,
                                     gate_vid=gate[:, :3072], gate_txt=gate[:, 3072:6144], gate_stride_b=6 * 3072))
        ms3 = t(lambda: dense.linear(a, w, b, out=out, epilogue=dense.EPI_GATE_RES))
        line += f" | gate_res {ms2:.3f} ms {fl/ms2/1e9:.0f} TF/s | gate_res(no gate) {ms3:.3f} ms"
    if N == 12288:
        ms1 = t(lambda: dense.linear(a, w, b, out=out, epilogue=dense.EPI_BIAS_GELU))
        line += f" | gelu {ms1:.3f} ms {fl/ms1/1e9:.0f} TF/s"
    print(line, flush=True)
