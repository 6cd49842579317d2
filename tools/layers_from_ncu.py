#!/usr/bin/env python
"""From one ncu launch list of `bench.py --steps 1 --warmup 3` (config 2; `--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --clock-control none --csv`) cut out ONE optimisation step (from one `adam_kernel` to the next) and write
  * the per-layer conv3x3 table (forward / dgrad / wgrad duration and TFLOP/s of all 17 layers, each launch alone) -> profiles/<tag>_layers.md
  * the step's DRAM traffic, whole and per kernel family, and that of the 17 forward launches -> profiles/traffic.json
Usage: tools/layers_from_ncu.py profiles/r02s_launches.csv r02s
The layer order is the model's: forward enc1b .. botb .. dec1b (enc1a is the CUDA-core first layer), backward the reverse; the dgrad of
dec1a is two row-streaming launches (csrc/igemm_conv3.cu: launch(), UB_CONV3_ROWS=3)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 16
FWD = [("enc1b", 64, 64, 512), ("enc2a", 64, 128, 256), ("enc2b", 128, 128, 256), ("enc3a", 128, 256, 128), ("enc3b", 256, 256, 128),
       ("enc4a", 256, 512, 64), ("enc4b", 512, 512, 64), ("bota", 512, 1024, 32), ("botb", 1024, 1024, 32), ("dec4a", 1024, 512, 64),
       ("dec4b", 512, 512, 64), ("dec3a", 512, 256, 128), ("dec3b", 256, 256, 128), ("dec2a", 256, 128, 256), ("dec2b", 128, 128, 256),
       ("dec1a", 128, 64, 512), ("dec1b", 64, 64, 512)]
BWD = [FWD[i] for i in (16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0)]


def flops(ci, co, hw):
    return 2.0 * 9 * ci * co * N * hw * hw


def load(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    L, order = {}, []
    for r in rows[1:]:
        i = int(r[ii])
        if i not in L:
            L[i] = {"name": r[ki]}
            order.append(i)
        L[i][r[mi]] = float(r[vi].replace(",", ""))
    return L, order


def kname(n):
    n = n.replace("void ", "").replace("<unnamed>::", "")
    return n[:n.index("(")] if "(" in n else n


def family(n):
    if "fold_conv3_kernel" in n:
        return "adam + repack + fold + small reductions"
    if "conv3_pair_kernel" in n or "conv3_rows_kernel" in n or "::conv3_kernel<" in n:
        return "conv3 fwd+dgrad"
    if "wgrad_halo" in n or "wgrad_sum_splits" in n:
        return "conv3 wgrad (+split sums)"
    if "igemm_fwd" in n or "igemm_wgrad" in n or "wgrad_reduce" in n:
        return "deconv fwd/dgrad/wgrad"
    if "bn_bwd_apply" in n:
        return "bn_bwd_apply"
    if "bn_bwd_reduce" in n:
        return "bn_bwd_reduce"
    if "pool_bwd" in n:
        return "pool_bwd_add"
    if "bn_apply" in n or "bn_pool" in n:
        return "bn_apply / bn_pool"
    if "head" in n:
        return "head"
    if "conv_first" in n:
        return "first layer"
    return "adam + repack + fold + small reductions"


def main():
    path, tag = sys.argv[1], sys.argv[2]
    L, order = load(path)
    adams = [i for i in order if "adam_kernel" in L[i]["name"]]
    assert len(adams) >= 2, "the capture must hold two Adam launches (one whole step between them)"
    a, b = adams[-2], adams[-1]
    step = [i for i in order if a < i <= b]
    us = lambda i: L[i]["gpu__time_duration.sum"] / 1e3
    byts = lambda i: L[i]["dram__bytes_read.sum"] + L[i]["dram__bytes_write.sum"]
    conv = [i for i in step if family(L[i]["name"]) == "conv3 fwd+dgrad"]
    wg = [i for i in step if "wgrad_halo_kernel" in L[i]["name"]]
    assert len(wg) == 17 and len(conv) in (34, 35), (len(wg), len(conv))
    split_dec1a = len(conv) == 35
    F = {n: (us(i), flops(ci, co, hw) / us(i) / 1e6, kname(L[i]["name"])) for i, (n, ci, co, hw) in zip(conv[:17], FWD)}
    D, j, dg = {}, 0, conv[17:]
    for n, ci, co, hw in BWD:
        k = 2 if (n == "dec1a" and split_dec1a) else 1
        t = sum(us(i) for i in dg[j:j + k])
        D[n] = (t, flops(ci, co, hw) / t / 1e6, " + ".join(kname(L[i]["name"]) for i in dg[j:j + k]))
        j += k
    W = {n: (us(i), flops(ci, co, hw) / us(i) / 1e6) for i, (n, ci, co, hw) in zip(wg, BWD)}
    tot = sum(flops(ci, co, hw) for _, ci, co, hw in FWD)
    out = [f"# Per-layer conv3x3 kernel times (config 2: 16x1x512^2, one step) from `{os.path.relpath(path, ROOT)}`", "",
           "Every launch alone (ncu serialises the streams) and cold, at the burst clock: upper bounds of the power-capped sustained rates.", "",
           "| layer | Cin -> Cout @ H | fwd us | fwd TFLOP/s | fwd kernel | dgrad us | dgrad TFLOP/s | dgrad kernel | wgrad us | wgrad TFLOP/s |",
           "|---|---|---:|---:|---|---:|---:|---|---:|---:|"]
    for n, ci, co, hw in FWD:
        f, d, w = F[n], D[n], W[n]
        out.append(f"| {n} | {ci} -> {co} @ {hw} | {f[0]:.1f} | {f[1]:.0f} | `{f[2]}` | {d[0]:.1f} | {d[1]:.0f} | `{d[2]}` | {w[0]:.1f} | {w[1]:.0f} |")
    tf, td, tw = sum(v[0] for v in F.values()), sum(v[0] for v in D.values()), sum(v[0] for v in W.values())
    out.append(f"| **all 17** | {tot / 1e12:.3f} TFLOP per pass | {tf:.0f} | {tot / tf / 1e6:.0f} | | {td:.0f} | {tot / td / 1e6:.0f} | | {tw:.0f} | {tot / tw / 1e6:.0f} |")
    with open(os.path.join(ROOT, "profiles", f"{tag}_layers_auto.md"), "w") as fh:
        fh.write("\n".join(out) + "\n")
    fam = {}
    for i in step:
        f = family(L[i]["name"])
        fam[f] = fam.get(f, 0.0) + byts(i) / 1e9
    rd = sum(L[i]["dram__bytes_read.sum"] for i in step)
    wr = sum(L[i]["dram__bytes_write.sum"] for i in step)
    frd = sum(L[i]["dram__bytes_read.sum"] for i in conv[:17])
    fwr = sum(L[i]["dram__bytes_write.sum"] for i in conv[:17])
    print(json.dumps({"step_launches": len(step), "step_ms_serialised": sum(us(i) for i in step) / 1e3, "step_dram_GB": (rd + wr) / 1e9,
                      "conv3_fwd_dram_GB": (frd + fwr) / 1e9, "conv3_fwd_us": tf, "by_family_GB": {k: round(v, 2) for k, v in fam.items()}}, indent=1))


if __name__ == "__main__":
    main()
