"""profiles/rNN_launches_summary.md from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
h = rows[0]
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    k = r[ik]
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", k)
    m = re.match(r"gemm_tcgen05_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(int\))?(\d+)>", k)
    k = f"gemm_tcgen05_kernel<CG={m.group(1)},BN={m.group(2)},EPI={m.group(3)}>" if m else re.sub(r"\(.*", "", k)[:90]
    tot[k] += float(r[iv].replace(",", "")) / 1000.0
    cnt[k] += 1
ENC = ("gemm_tcgen05", "attention", "layernorm", "preprocess", "cls_rows", "gather_rows")
all_us = sum(tot.values())
enc_us = sum(v for k, v in tot.items() if k.startswith(ENC) or "attention" in k)
gemm_us = sum(v for k, v in tot.items() if k.startswith("gemm_tcgen05"))
with open(dst, "w") as f:
    f.write(f"# ncu launch list, `{cmd}`\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (setup + warm-up + timed steps). "
            "Per-launch times are cold-cache and serialised: compare SHARES.\n\n")
    f.write("| kernel | launches | total us | avg us | share of all | share of encoder kernels |\n|---|---|---|---|---|---|\n")
    for k, v in tot.most_common(24):
        enc = f"{v / enc_us * 100:.1f}%" if (k.startswith(ENC) or "attention" in k) else "-"
        f.write(f"| `{k}` | {cnt[k]} | {v:.1f} | {v / cnt[k]:.1f} | {v / all_us * 100:.1f}% | {enc} |\n")
    f.write(f"\nGEMM share of the encoder kernels under ncu: {gemm_us / enc_us * 100:.1f}% "
            "(bench.py live CUDA-event share of the step: see BENCH json `roofline.kernel_share_of_step`).\n")
print(open(dst).read())
