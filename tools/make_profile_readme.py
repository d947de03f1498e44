#!/usr/bin/env python
"""Builds profiles/README.md and profiles/ncu_traffic.json from the measurement files gathered on the B200 box."""
import csv
import json
import re
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"
R = "r01"


def main():
    bench = json.loads((P / f"{R}_bench.json").read_text())
    prof = json.loads((P / f"{R}_launch_profile.json").read_text())
    aux = [json.loads(l) for l in (P / f"{R}_aux_kernels.jsonl").read_text().splitlines() if l.strip()]
    ncu = list(csv.reader((P / f"{R}_kernels_ncu_full.csv").open()))
    out = []
    w = out.append
    w(f"# profiles — round 1 (B200, sm_100a)\n")
    w("All numbers were produced on the pool's B200 boxes through `gpurun`. Timed numbers come from CUDA events "
      "(bench.py / tools/*.py); ncu numbers are cold-cache, serialised replays and are used for SHARES and DRAM "
      "traffic only.\n")
    w("## Headline (`r01_bench.json` = `python bench.py --steps 40 --warmup 3`)\n")
    rf, e2e, cpu, ck = bench["roofline"], bench["e2e"], bench["cpu_baseline"], bench["clocks"]
    w(f"| quantity | value |\n|---|---|")
    w(f"| workload | {bench['config']['workload']}, {bench['config']['pairs_per_step']} pairs per step |")
    w(f"| `value` (device-timed, inputs in HBM) | **{bench['value']:.1f} frames/s** ({bench['ms_per_step']:.2f} ms per step) |")
    w(f"| `e2e` (host u8 clip -> host u8 frames through `fiNetInterpolateClipHostU8`) | **{e2e['value']:.1f} frames/s** |")
    w(f"| whole-step arithmetic rate | {rf['whole_step_tflops']:.0f} TFLOP/s |")
    w(f"| tcgen05 conv launches (96.9 % of the step) | {rf['achieved']:.0f} TFLOP/s = **{rf['frac']*100:.1f} %** of the measured sustained bf16 peak ({rf['peak']:.0f}) |")
    w(f"| clocks during the timed region | {ck['sm_mhz']:.0f} MHz median of {ck['sm_max_mhz']:.0f}, reasons {ck['reasons']} |")
    w(f"| CPU baseline (oracle port, {cpu['cores']} host cores) | {cpu['value']:.3f} frames/s |")
    w("")
    sc = P / f"{R}_scaling.jsonl"
    if sc.exists():
        rows = [json.loads(l) for l in sc.read_text().splitlines() if l.strip()]
        w("## Scaling (`r01_scaling.jsonl`: `torchrun --nproc-per-node N bench.py --gpus N --steps 20 --warmup 3`, "
          "frame pairs sharded by rank, no collective on the data path)\n")
        w("| GPUs | frames/s (device-timed, max over ranks) | e2e frames/s (host clip in, host frames out) | x of 1 GPU |\n|---|---|---|---|")
        for r in rows:
            w(f"| {r['n_gpus']} | {r['value']:.0f} | {r['e2e']:.0f} | {r['value']/rows[0]['value']:.2f} |")
        w("")
    w("## Per-launch table (`r01_launch_profile.json`, CUDA events inside the timed region, 4 pairs per launch)\n")
    w("| launch | kernel | ms | TFLOP/s | algorithmic GB/s | share |\n|---|---|---|---|---|---|")
    tot = sum(p["ms_total"] for p in prof)
    for p in prof:
        ms = p["ms_total"] / p["calls"]
        kind = {0: "stem_mma", 1: "tcgen05 conv", 2: "upsample"}[p["kind"]]
        w(f"| {p['name']} | {kind} | {ms:.3f} | {p['flops']/ms/1e9:.0f} | {p['bytes']/ms/1e6:.0f} | {p['ms_total']/tot*100:.1f} % |")
    w(f"| **total** | | **{tot/prof[0]['calls']:.2f}** | | | |")
    w("")
    w("## ncu (`r01_kernels_ncu_full.csv`: `--set full`, one forward at 1 pair; `r01_launches_ncu.csv`: launch list)\n")
    h = ncu[0]
    col = {name: i for i, name in enumerate(h)}

    def find(prefix):
        return [i for i, n in enumerate(h) if n.startswith(prefix)][0]
    ik, it, ir, iw = find("Kernel Name"), find("gpu__time_duration"), find("dram__bytes_read"), find("dram__bytes_write")
    itn = find("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    il2 = find("lts__t_sector_hit_rate")
    w("| # | kernel | us | DRAM read MB | DRAM write MB | tensor pipe active % | L2 hit % |\n|---|---|---|---|---|---|---|")
    dram = 0.0
    n_conv = 0
    for i, r in enumerate(ncu[1:]):
        m = re.search(r"(\w+_kernel<[^>]*>)", r[ik])
        name = m.group(1) if m else r[ik][:60]
        w(f"| {i} | `{name}` | {float(r[it]):.1f} | {float(r[ir]):.1f} | {float(r[iw]):.1f} | {float(r[itn]):.1f} | {float(r[il2]):.1f} |")
        if "stem" not in name:
            dram += (float(r[ir]) + float(r[iw])) * 1e6
            n_conv += 1
    w("")
    algo = sum(p["bytes"] for p in prof if p["kind"] == 1) / bench["config"]["pairs_per_step"]
    w(f"DRAM traffic of the {n_conv} tcgen05 conv launches of one forward (1 pair): **{dram/1e9:.2f} GB** measured vs "
      f"{algo/1e9:.2f} GB algorithmic (every activation/weight touched once): no re-read inflation — the 9 taps and the "
      "halo overlap are served from L2/SMEM.\n")
    (P / "ncu_traffic.json").write_text(json.dumps({"source": f"{R}_kernels_ncu_full.csv", "pairs": 1,
                                                    "conv_launches": n_conv, "dram_bytes": dram,
                                                    "algorithmic_bytes": algo}, indent=1))
    w("## Non-GEMM kernels (`r01_aux_kernels.jsonl` = `python tools/bench_aux.py`)\n")
    w("A write-only stream reaches 3.9 TB/s on this GPU against 6.45 TB/s for a 1:1 copy (`r01_bw_probe.txt`), so a "
      "kernel reading R and writing W bytes is bounded by max((R+W)/copy peak, W/3.9 TB/s); the last column is the "
      "measured time against that bound.\n")
    w("| kernel | ms | achieved GB/s | of HBM copy peak | write share | of the read/write-mix bound | note |\n|---|---|---|---|---|---|---|")
    write_share = {"pack_pair_u8": 8 / 10, "head_post_u8": 1 / 5, "upsample2x_bilinear": 4 / 5,
                   "stem_conv (tcgen05, hi/lo split)": 128 / 130}
    for a in aux:
        ws = write_share.get(a["kernel"], 0.0)
        total = a["algorithmic_bytes"]
        bound_ms = max(total / (a["hbm_peak_gbs"] * 1e6), ws * total / (3900.0 * 1e6))
        w(f"| {a['kernel']} | {a['ms']} | {a['achieved_gbs']} | {a['frac_of_hbm_peak']*100:.1f} % | {ws*100:.0f} % | "
          f"{bound_ms / a['ms'] * 100:.0f} % | {a['note']} |")
    w("")
    cfg = P / f"{R}_configs.jsonl"
    if cfg.exists():
        w("## Other BASELINE configs (`r01_configs.jsonl` = `python tests/bench_configs.py`)\n")
        for l in cfg.read_text().splitlines():
            if l.strip():
                w("```json\n" + l + "\n```")
        w("")
    tb = P / f"{R}_train_bench.jsonl"
    if tb.exists():
        rows = [json.loads(l) for l in tb.read_text().splitlines() if l.strip().startswith("{")]
        w("## Training step (`r01_train_bench.jsonl` = `python tools/bench_train.py [--graph] [--criterion combined]`)\n")
        w("One optimisation step (train-mode forward, loss, backward, Adam) of `FrameInterpolationUNet(bilinear=True)`, "
          "CUDA events around 20 steps after 5 warm-up steps; the torch arms run the same network as eager torch ops "
          "on the same GPU.\n")
        w("| arm | workload | ms / step | samples/s |\n|---|---|---|---|")
        for r in rows:
            w(f"| {r['arm']}{' x' + str(r['n_gpus']) + ' GPUs' if r.get('n_gpus', 1) > 1 else ''} | "
              f"{r['workload'].replace('train step: FrameInterpolationUNet(bilinear) ', '')} | "
              f"{r['ms_per_step']:.2f} | {r['samples_per_s']:.0f} |")
        w("")
        tr = P / f"{R}_train_trace.txt"
        if tr.exists():
            w("Kernel timeline of one replayed step (`r01_train_trace.txt` = `python tools/trace_train.py --graph "
              "--no-overlap`, CUPTI, warm): conv forward + data gradient ~1.8 ms and weight gradient ~1.6 ms on the "
              "tensor cores (4.6 TFLOP per step), the four BatchNorm passes ~1.8 ms at ~4.4 TB/s, everything else "
              "~1.2 ms; no idle gaps. `r01_train_launches_ncu.csv` is the ncu launch list of an eager step.\n")
            w("```\n" + tr.read_text().strip() + "\n```\n")
    (P / "README.md").write_text("\n".join(out) + "\n")
    print("\n".join(out)[:3000])


if __name__ == "__main__":
    main()
