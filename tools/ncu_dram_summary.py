"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of ONE step
(tools/one_step.py) per kernel family: launches, device time, DRAM bytes (read + write, per launch and total), achieved GB/s
against the measured HBM peak.  usage: ncu_dram_summary.py step_dram.csv out.json [source-label]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = [("gemm", r"gemm_kernel|gemm_multi_kernel"), ("attn_teacher_fwd", r"attn_fwd_tc_kernel"), ("attn_student_fwd", r"attn_fwd_lse"),
            ("attn_student_bwd", r"attn_bwd"), ("attn_mma_sync", r"flash_"), ("cls_attn", r"cls_attn"),
            ("layernorm_fwd", r"layernorm_fwd|ln_fwd"), ("layernorm_bwd", r"layernorm_bwd|ln_bwd"), ("teacher_embed_ln", r"teacher_embed"),
            ("dec_tail_fwd", r"dec_tail_fwd"), ("dec_tail_bwd", r"dec_tail_bwd"), ("l2norm_rows", r"l2norm"), ("patchify", r"patchify"),
            ("gather_rows", r"gather_rows"), ("colsum_bf16", r"colsum"), ("cast_scale", r"cast_scale|cast_bf16"), ("mask_select", r"mask_select"),
            ("adamw", r"adamw"), ("sumsq", r"sumsq"), ("drop_path_draw", r"drop_path"), ("torch (fill / copy / elementwise)", r"at::|cub::")]


def main():
    src, out = sys.argv[1], sys.argv[2]
    label = sys.argv[3] if len(sys.argv) > 3 else src
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = {}
    for r in rows[1:]:
        d = per.setdefault(r[iid], dict(name=r[ik]))
        d[r[im]] = float(r[iv].replace(",", ""))
    hbm = 6529.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = json.load(open(p)).get("hbm_gbs", hbm)
    fam = {}
    for d in per.values():
        name = next((f for f, pat in FAMILIES if re.search(pat, d["name"])), "other: " + d["name"][:60])
        a = fam.setdefault(name, dict(launches=0, ns=0.0, dram_read=0.0, dram_write=0.0))
        a["launches"] += 1
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
        a["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
    total_ns = sum(a["ns"] for a in fam.values())
    res = {}
    for name, a in sorted(fam.items(), key=lambda kv: -kv[1]["ns"]):
        b = a["dram_read"] + a["dram_write"]
        res[name] = dict(launches=a["launches"], ms=round(a["ns"] / 1e6, 4), share=round(a["ns"] / total_ns, 4),
                         us_per_launch=round(a["ns"] / 1e3 / a["launches"], 2), dram_bytes=int(b), dram_read=int(a["dram_read"]),
                         dram_write=int(a["dram_write"]), dram_mb_per_launch=round(b / a["launches"] / 1e6, 2),
                         gb_per_s=round(b / max(a["ns"], 1.0), 1), frac_of_measured_hbm=round(b / max(a["ns"], 1.0) / hbm, 3))
    json.dump(dict(source=label, note="per-launch times under ncu are cold-cache and serialised: compare shares; DRAM bytes are exact",
                   hbm_peak_gbs=hbm, total_kernel_ms=round(total_ns / 1e6, 3), families=res), open(out, "w"), indent=1)
    for k, v in res.items():
        print(f"{k:36s} n={v['launches']:4d} {v['ms']:8.3f} ms {v['share']*100:5.1f}%  {v['dram_mb_per_launch']:9.2f} MB/launch {v['gb_per_s']:8.1f} GB/s ({v['frac_of_measured_hbm']:.2f})")


if __name__ == "__main__":
    main()
