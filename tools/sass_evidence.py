"""Dev: per-kernel counts of the Blackwell-native SASS mnemonics in the built library -> profiles/<round>_sass_evidence.md
usage: python tools/sass_evidence.py r02"""
import collections, re, subprocess, sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", "videogpa_b200/lib/libvideogpa_b200.so"], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "FMUL2"]
cur, stats, samples = None, collections.OrderedDict(), {}
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); stats[cur] = collections.Counter(); samples[cur] = {}
        continue
    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if cur is None or not mm:
        continue
    ins = mm.group(2)
    stats[cur]["_total"] += 1
    for k in keys:
        if re.search(r"(^|\s)" + re.escape(k), ins):
            stats[cur][k] += 1
            samples[cur].setdefault(k, [])
            if len(samples[cur][k]) < 1:
                samples[cur][k].append(f"/*{mm.group(1)}*/ {ins.strip()}")
names = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
rows = []
for (k, c), n in zip(stats.items(), names):
    if c["UTCHMMA"] or c["UTMALDG"] or c["FFMA2"] or c["LDTM"]:
        n = n.replace("(anonymous namespace)::", "").replace("vgpa::", "").replace("void ", "").replace("(int)", "")
        n = re.sub(r"\(.*", "", n)
        rows.append((n, c, k))
rows.sort(key=lambda r: r[0])
out = [f"# {rnd} SASS evidence: Blackwell-native instructions per kernel of `videogpa_b200/lib/libvideogpa_b200.so`", "",
       "`cuobjdump -sass` of the library built by `python -m videogpa_b200.build` (sm_100a); static instruction counts per kernel.",
       "UTCHMMA = tcgen05.mma (bf16), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier",
       "ops, FFMA2 / FADD2 / FMUL2 = packed fp32x2 FMA-pipe ops. Regenerate with `python tools/sass_evidence.py " + rnd + "`.", "",
       "| kernel | SASS instr | UTCHMMA | LDTM | STTM | UTMALDG | UTMASTG | UTCBAR | SYNCS | MUFU.EX2 | FFMA2 | FADD2 | FMUL2 |", "|" + "---|" * 13]
for n, c, k in rows:
    out.append(f"| `{n}` | {c['_total']} | " + " | ".join(str(c[x]) for x in keys) + " |")
out += ["", "## Sample instructions (first occurrence per kernel)", "", "```"]
for n, c, k in rows:
    if any(x in n for x in ["attn_fwd_d64_bounded_kernel<4>", "gemm_bf16_kernel<256, 1, 2>", "attn_fwd_d128", "vae_conv3d_kernel<128, 0", "attn_bwd_dq", "attn_bwd_dkv",
                            "mvcs_pairs"]):
        out.append(f"--- {n}")
        for kk in ["UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "FFMA2", "MUFU.EX2"]:
            for smp in samples[k].get(kk, []):
                out.append("    " + smp[:160])
out += ["```", "", "No kernel uses `UTMASTG` (TMA store): the epilogues store 16-byte vectors from registers after the fused per-row arithmetic."]
open(f"profiles/{rnd}_sass_evidence.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
