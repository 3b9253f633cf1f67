#!/usr/bin/env python
"""Summarise ncu --set full captures (no GPU needed): per captured launch the duration, DRAM bytes, tensor-pipe / issue activity, and the
per-stage DRAM traffic table bench.py reads (profiles/ncu_traffic.json).

  python scripts/ncu_summary.py gpurun_out/r02_final.ncu-rep [gpurun_out/r02_ctcfc.ncu-rep ...]
writes profiles/r02_ncu_full_top_kernels.md, profiles/r02_ncu_raw.csv, profiles/ncu_traffic.json
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread"]
I = r"(\(int\))?"
STAGE_OF = [(r"stft_power_warp_kernel|db_mel_fast_kernel", "stft"), (r"conv1_f32x2_kernel|conv1_kernel", "conv1"),
            (r"gemm_tc_kernel<" + I + r"1,", "conv2"), (r"conv_subsample_tc_kernel", "conv2_fused"), (r"gemm_tc_kernel<" + I + r"8,", "sub_linear"),
            (r"gemm_chain_pair_kernel<" + I + r"6, (\(bool\))?(0|false)>", "ffn_chain"), (r"gemm_tc_kernel<" + I + r"5,", "qkv"),
            (r"attention_tc_kernel", "attention"), (r"gemm_chain_pair_kernel<" + I + r"6, (\(bool\))?(1|true)>", "out_proj"),
            (r"gemm_tc_kernel<" + I + r"3,", "pw1_glu"), (r"dwconv_reg", "dwconv"), (r"gemm_tc_kernel<" + I + r"9,", "ctc_fc"),
            (r"gemm_tc_kernel<" + I + r"0,", "ctc_fc_logits")]


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)


def main():
    reps = sys.argv[1:]
    rows_out = []
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            d = {"report": os.path.basename(rep), "id": r[ix["ID"]], "kernel": r[ix["Kernel Name"]]}
            for c in COLS:
                if c in ix:
                    try:
                        d[c] = float(r[ix[c]]) * (unit_scale(units[ix[c]]) if ("bytes" in c or "duration" in c) else 1.0)
                    except ValueError:
                        d[c] = None
            rows_out.append(d)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r02_ncu_raw.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["report", "id", "kernel"] + COLS)
        for d in rows_out:
            w.writerow([d["report"], d["id"], d["kernel"]] + [d.get(c) for c in COLS])
    traffic, seen = {}, {}
    md = ["# Round 2: `ncu --set full --clock-control none` of the step's kernels (BASELINE config 2, 32 x 10 s, one B200)", "",
          "Raw metric columns: `profiles/r02_ncu_raw.csv` (written by `scripts/ncu_summary.py` from the `.ncu-rep` files; cold-cache, serialised",
          "launches: durations are NOT bench numbers, the DRAM bytes and pipe activities are the evidence).", "",
          "| # | kernel | stage | duration us | DRAM read MB | DRAM write MB | tensor pipe active % | issue slots busy % | L2 hit % | grid x block | regs |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    for d in rows_out:
        stage = next((s for rx, s in STAGE_OF if re.search(rx, d["kernel"])), "")
        name = re.sub(r"\(CUtensorMap_st.*|\(const.*|\(b200asr.*|\(Attn.*|\(Dw.*|\(Conv.*", "", d["kernel"]).replace("b200asr::<unnamed>::", "").replace("void ", "")
        rd, wr = (d.get("dram__bytes_read.sum") or 0) / 1e6, (d.get("dram__bytes_write.sum") or 0) / 1e6
        md.append(f"| {d['id']} | `{name[:60]}` | {stage} | {d.get('gpu__time_duration.sum', 0):.1f} | {rd:.2f} | {wr:.2f} | "
                  f"{d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed') or 0:.1f} | "
                  f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active') or 0:.1f} | {d.get('lts__t_sector_hit_rate.pct') or 0:.0f} | "
                  f"{int(d.get('launch__grid_size') or 0)} x {int(d.get('launch__block_size') or 0)} | {int(d.get('launch__registers_per_thread') or 0)} |")
        if stage and stage not in seen:
            seen[stage] = True
            traffic[stage] = traffic.get(stage, 0) + (rd + wr) * 1e6
        elif stage == "stft" and seen.get("stft") is True:       # stft = two kernels (STFT + dB/mel): add the second one once
            traffic["stft"] += (rd + wr) * 1e6
            seen["stft"] = 2
    md += ["", "Per-stage DRAM bytes of ONE launch (first occurrence in the capture) -> `profiles/ncu_traffic.json`, read by `bench.py` for",
           "`roofline.traffic` / `other_stages.*.traffic`."]
    open(os.path.join(ROOT, "profiles", "r02_ncu_full_top_kernels.md"), "w").write("\n".join(md) + "\n")
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    merged = json.load(open(tpath)) if os.path.isfile(tpath) else {}          # stages this capture does not cover keep their earlier value
    merged.update({k: round(v) for k, v in traffic.items()})
    json.dump(merged, open(tpath, "w"), indent=1)
    print("\n".join(md))
    print(traffic)


if __name__ == "__main__":
    main()
