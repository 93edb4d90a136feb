#!/usr/bin/env python
"""Turn what a GPU session (tools/gpu_session.sh <tag>, tools/gpu_session_multi.sh <mtag>) left in gpurun_out/
into the tracked summaries under profiles/.  Runs in the build container (no GPU; `ncu -i` reads the report).

    python tools/make_profiles.py --tag r02p --multi r02m8 --out r02
"""
import argparse
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

NCU_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu_variants(tag, out):
    res = {"note": "ncu --set full --clock-control none, caches flushed before every replay (= production regime), one launch "
                   "per variant of tools/ncu_target.py <case> on the 25,124,842-parameter chain; final code of the round",
           "variants": {}}
    for case, f in (("SGLD.step (Philox, register path)", "step_cold"),
                    ("VerletSGLD.step + fused prior (TMA-staged)", "verlet_fused_cold"),
                    ("SGLD.step(calc_metrics=True) (all sums, TMA-staged)", "sgld_metrics_cold"),
                    ("HMC.step (no noise)", "hmc_cold"), ("VerletSGLD.initial_step(save_state)", "verlet_save_cold")):
        path = os.path.join(G, f"{tag}_{f}_raw.csv")
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        r, u = dict(zip(hdr, rows[2])), dict(zip(hdr, units))
        res["variants"][case] = {"kernel": r["Kernel Name"],
                                 **{k: {"value": float(r[k].replace(",", "")), "unit": u[k]} for k in NCU_KEYS if k in r}}
    json.dump(res, open(os.path.join(P, f"{out}_ncu_variants.json"), "w"), indent=1)


def parity_md(tag, out):
    path = os.path.join(G, f"{tag}_parity_reports.jsonl")
    if not os.path.exists(path):
        return
    rows = [json.loads(l) for l in open(path)]
    md = ["# Parity on the BASELINE configs (GPU, `gpurun`, final code of the round)", "",
          "## Under the reference's own runners (`tests/test_cuda_reference_runners.py`)", "",
          "Run A = reference runner + reference eager sampler on cuda:0 (recorded); run B = same runner after `overlay.install()` "
          "(the kernel), every sampler call fed run A's inputs and compared with run A's outputs. Errors: worst over all calls; "
          "`p`/`m` relative to the tensor's RMS; ΔE against the size of its terms; decisions = Metropolis tests (equal / total, rejections); "
          "kink = elements a fused replay found within 1e-5 of the reference's value but on the other side of zero, put on the reference's "
          "value before the step (a Laplace prior's gradient term jumps there; tests/runner_tape.py).", "",
          "| case | sampler calls | p | m | est_temperature | est_config_temp | ΔE (terms) | max abs ΔE | decisions equal | rejections | kink |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        if r.get("golden"):
            continue
        se = r["scalar_err"]
        md.append(f'| {r["case"]} | {r["n_events"]} | {r["p_err"]:.1e} | {r["m_err"]:.1e} | {se.get("est_temperature", 0):.1e} | '
                  f'{se.get("est_config_temp", 0):.1e} | {r["de_term_err"]:.1e} | {r["de_scale"]:.0f} | {r["decisions_equal"]}/{r["decisions"]} | {r["rejections"]} | {r.get("kink_guards", 0)} |')
    md += ["", "## Compact goldens at the real segment tables (`tests/test_cuda_real_tables.py`)", "",
           "Reference sampler on CPU (recorded as seeds + fingerprints) vs the CUDA sampler; `+fused_prior` = prior evaluated by the kernel.", "",
           "| case | events | strided samples | moments | ΔE (terms) | decisions equal | rejections | worst scalar |", "|---|---|---|---|---|---|---|---|"]
    for r in rows:
        if not r.get("golden"):
            continue
        ws = max(r["scalar_err"].items(), key=lambda kv: kv[1]) if r["scalar_err"] else ("-", 0.0)
        md.append(f'| {r["case"]} | {r["n_events"]} | {r["sample_err"]:.1e} | {r["moment_err"]:.1e} | {r["de_term_err"]:.1e} | '
                  f'{r["decisions_equal"]}/{r["decisions"]} | {r["rejections"]} | {ws[0]} {ws[1]:.1e} |')
    open(os.path.join(P, f"{out}_runner_parity.md"), "w").write("\n".join(md) + "\n")


def notes_md(out, have_multi):
    b = json.load(open(os.path.join(P, f"{out}_bench.json")))
    b20 = json.load(open(os.path.join(P, f"{out}_bench_20steps.json")))
    ref = json.load(open(os.path.join(P, f"{out}_bench_reference_arm.json")))
    sm = json.load(open(os.path.join(P, f"{out}_small_models.json")))
    tune = json.load(open(os.path.join(P, f"{out}_tune.json")))
    steady = json.load(open(os.path.join(P, f"{out}_ncu_dram_steady.json")))
    cold = json.load(open(os.path.join(P, "ncu_step_kernel.json")))
    r = b["roofline"]
    L = [f"# Round 2 — measured numbers (B200 via `gpurun`, final code, SM clock {b['clocks']['sm_mhz']} MHz, clock reasons: {b['clocks']['reasons'] or 'none'})\n",
         "Every row names the file and the command behind it; bench lines are `bench.py` output verbatim. Generated by `tools/make_profiles.py`.\n",
         f"## Headline workload: 25,124,842-parameter chain (54 tensors), SGLD, N=1 (`{out}_bench.json`: `python bench.py --steps 200 --warmup 5`)\n",
         "| what | value |\n|---|---|",
         f"| `value` (API `opt.step(calc_metrics=False)`) | {b['value']:.4g} param-updates/s, {b['ms_per_step']*1e3:.2f} µs/step, {b['gpu_launches']} launches in {b['steps']} steps |",
         f"| same with 20 timed steps (`{out}_bench_20steps.json`) | {b20['value']:.4g}, {b20['ms_per_step']*1e3:.2f} µs/step (kernel {b20['roofline']['kernel_us']:.2f}) |",
         f"| `roofline.frac` (back to back, alternating direction) | {r['frac']:.3f} = {r['achieved']:.0f} GB/s of {r['peak']}; kernel {r['kernel_us']:.2f} µs; DRAM traffic {steady['dram_bytes_per_launch']/1e6:.1f} MB/launch (`{out}_ncu_dram_steady.json`) = {r['dram_GBs']:.0f} GB/s = {r['dram_frac']:.2f} of the copy peak |"]
    pr, sd = r["production"], r["same_direction"]
    L.append(f"| **`roofline.production`** (L2 evicted before every launch) | **{pr['kernel_us']:.2f} µs, frac {pr['frac']:.3f}; 20 B/param touched = {pr['touched_GBs']:.0f} GB/s = {pr['touched_frac']:.3f} of the copy peak**; ncu cold capture {cold['dram_bytes_per_launch_cold']/1e6:.1f} MB while the kernel runs (`{out}_ncu_step_kernel.json`) |")
    L.append(f"| `roofline.same_direction` | {sd['kernel_us']:.2f} µs, frac {sd['frac']:.3f}, touched {sd['touched_frac']:.3f} |")
    for k, v in b["extra"].items():
        if "what" in v:
            L.append(f"| API step in the reference runner's loop (`zero_grad()`, gradients in tensors of their own, L2 evicted) | {v['us_per_step']:.2f} µs, frac {v['frac']:.3f}, gradient copies {v['gradient_copies']}, pointer-table writes {v['pointer_table_writes']} (round 1: ≈110 µs with the 8 B/param copy) |")
    for k, v in b["samplers"].items():
        L.append(f"| {k} (all ranks) | {v['value']:.4g} param-updates/s; kernel {v['kernel_us']:.2f} µs (frac {v['frac']:.3f}); production {v['production_kernel_us']:.2f} µs (frac {v['production_frac']:.3f}) |")
    for k, v in b["extra"].items():
        if "production_us_per_step" in v:
            L.append(f"| {k} | {v['us_per_step']:.2f} µs back to back, {v['production_us_per_step']:.2f} µs production |")
        elif "us_per_step" in v and "what" not in v:
            L.append(f"| {k} | {v['us_per_step']:.2f} µs (API, back to back) |")
    e = b["e2e"]
    L += [f"| `e2e` (gradient from pinned host memory in, diagnostics out, every step) | {e['value']:.4g} param-updates/s, {e['ms_per_step']:.3f} ms/step; the bare copy alone: {e['bare_h2d_ms']:.3f} ms = {e['bare_h2d_GBs_per_gpu']:.1f} GB/s |",
          f"| `cpu_baseline` (numpy port, 1 thread) | {b['cpu_baseline']['value']:.3g} |",
          f"| host time per `opt.step` in the bench loop | {b['impl_notes']['host_us_per_step']:.1f} µs |",
          f"| other kernels (`{out}_tune.json`, `tools/tune_tiles.py`) | rollback {tune['rollback'][1]} µs (603 MB), bare access pattern `bnnp_probe_stream` {tune['probe_stream'][1]} µs, step with 21 sampled scales {tune['verlet_hier'][1]} µs (round 1 / before the chained epilogue: 93), pre-pass + epilogue {tune['hier_prepass_plus_epilogue'][1]} µs |",
          f"\n## Reference arm (`{out}_bench_reference_arm.json`: `python bench.py --impl reference --steps 20 --warmup 5`, {ref['cpu_baseline']['cores']} host cores)\n",
          "| arm | kind | ms/step | param-updates/s |\n|---|---|---|---|"]
    for k, v in ref["cpu_arms"].items():
        L.append(f"| {k} | {v['kind']} | {v['ms_per_step']:.1f} | {v['value']:.3g} |")
    if ref.get("reference_calls"):
        L.append("\nThe unmodified reference classes (`oracle/_ref`, BASELINE.md §4), ms per call: " +
                 ", ".join(f"`{k}` {v['ms_per_step']:.1f}" for k, v in ref["reference_calls"].items()) + ".\n")
    if have_multi:
        ns = {n: json.load(open(os.path.join(P, f"{out}_bench_n{n}.json"))) for n in (1, 2, 4, 8)
              if os.path.exists(os.path.join(P, f"{out}_bench_n{n}.json"))}
        L += [f"## 1 → 8 GPUs (`{out}_bench_n*.json`: torchrun, `--steps 20 --warmup 5 --no-extra --no-cpu`, one 8-GPU box)\n",
              "| N | SGLD value | ms/step − kernel µs | efficiency | VerletSGLD | HMC | e2e | bare H2D GB/s per GPU | cycle gather ms |\n|---|---|---|---|---|---|---|---|---|"]
        v1 = ns[1]["value"]
        for n, d in ns.items():
            L.append(f"| {n} | {d['value']:.4g} | {d['ms_per_step']*1e3 - d['roofline']['kernel_us']:+.2f} | {d['value']/(n*v1):.3f} | {d['samplers']['VerletSGLD.step']['value']:.4g} | "
                     f"{d['samplers']['HMC.step']['value']:.4g} | {d['e2e']['value']:.4g} | {d['e2e']['bare_h2d_GBs_per_gpu']:.1f} | {d.get('cycle_gather_ms') and round(d['cycle_gather_ms'], 3)} |")
        L.append("\n`smoke()` on the 2-GPU and on the 8-GPU box: `N chains on N GPUs, NCCL gather == independent runs (bit for bit)`; "
                 "`tests/test_cuda_chains_nccl.py` passed on both (final code of the round).\n")
    L += [f"## Small BASELINE configs (`{out}_small_models.json`: `python tools/bench_small_models.py`), µs per step\n",
          "| config | tensors | host, views | host, runner loop (`zero_grad` / `step`) | kernel | graph replay host / wall | round-1 host |\n|---|---|---|---|---|---|---|"]
    r1 = json.load(open(os.path.join(P, "r01g_small_models.json")))
    for k, v in sm.items():
        L.append(f"| {k} | {v['tensors']} | {v['api_host_us_per_step']} | {v['runner_loop_zero_grad_host_us']} / {v['runner_loop_step_host_us']} | {v['kernel_back_to_back_us']} | "
                 f"{v['graph_replay_host_us']} / {v['graph_replay_wall_us']} | {r1.get(k, {}).get('api_host_us_per_step', '-')} |")
    names = {"a_base.so": "round-2 code before the split", "e_bm_only.so": "lean Box-Muller", "f_full2_bm.so": "lean B-M + full-chunk path for all-sums variants",
             "g_full2_nobm.so": "full-chunk path for all-sums only", "h_none.so": "neither", "default": "FINAL library",
             "a_tma0.so": "register path (no TMA)", "b_tma2.so": "TMA staging for fused-prior + all-sums variants", "c_tma1.so": "TMA staging everywhere",
             "d_tma2_c4.so": "the same as column 2 + all-sums variants at 4 CTAs/SM (kept)"}
    for title, fname in (("Kernel variants I: full-chunk code path and lean Box–Muller", f"{out}_variants.jsonl"),
                         ("Kernel variants II: TMA-staged chunks (cp.async.bulk + mbarrier)", f"{out}_variants_tma.jsonl"),
                         ("Kernel variants: the final library", f"{out}_variants_final.jsonl")):
        path = os.path.join(P, fname)
        if not os.path.exists(path):
            continue
        rows = [json.loads(l) for l in open(path) if l.strip()]
        cases = [k for k in rows[0] if k != 'lib']
        L.append(f"\n## {title}, back to back / production µs (`{fname}`: `tools/variant_times.py` per build)\n")
        L.append("| case | " + " | ".join(names.get(r_['lib'], r_['lib']) for r_ in rows) + " |\n|---|" + "---|" * len(rows))
        for c in cases:
            L.append(f"| {c} | " + " | ".join(f"{r_[c]['b2b_us']} / {r_[c]['production_us']}" if c in r_ else "-" for r_ in rows) + " |")
    L.append("\nRejected on the way (session logs, not kept): a full-chunk register path for EVERY variant made `verlet_fused` slower "
             "(82.7 → 87.1 µs production, more spills in the eight per-form copies); `BNNP_ONLY_NORMAL` (upper bound of a common-forms-only "
             "instantiation) 85.6 µs there; 3 CTAs/SM for the prior variants 89.3 µs; removing the control-block branch or the gradient "
             "pointer indirection changed nothing (±0.5 µs). The GPU test-suite also passes with the staged path forced on for every variant "
             "(`BNNP_TMA_MODE=1` build, 157 kernel-level tests).\n")
    wpath = os.path.join(P, f"{out}_runner_wallclock.json")
    if os.path.exists(wpath):
        w = json.load(open(wpath))
        L.append(f"## The reference's own `experiments/train_bnn.py`, unmodified, wall-clock on one B200 (`{out}_runner_wallclock.json`: `python tools/bench_runner.py`)\n")
        L.append("Synthetic data of the data set's shape, 6,400 training points at batch 128 (50 minibatches per epoch = 50 leapfrog steps for HMC), 2 cycles × (1 warm-up + 2 sampling epochs) "
                 "= 300 sampler steps, 4 Metropolis tests, per-epoch evaluation on 1,024 test points, samples written to the HDF5 stand-in. Second run of each (the first pays cuDNN autotuning).\n")
        L.append("| config | stock reference | `overlay.install()` (sampler only) | `overlay.install(evaluate=True, fuse_prior=True, sample_sink=True)` | speed-up |\n|---|---|---|---|---|")
        for k, v in w.items():
            L.append(f"| {k} | {v['reference']['seconds']} s | {v['overlay']['seconds']} s | {v['overlay_all']['seconds']} s | {v['speedup_overlay']}× / {v['speedup_overlay_all']}× |")
        L.append("\nWhat remains is the reference's own loop: forward / backward passes (incl. the full-data gradient pass of every sampling epoch), its DataLoader, `.item()` reads and metric bookkeeping.\n")
    L.append("## ncu\n")
    L.append(f"* `{out}_ncu_step_kernel.json` — `--set full` of `bnnp_step_kernel<2,0,0,0>` (SGLD, Philox), caches flushed before every replay: DRAM bytes while the kernel runs, registers, occupancy, stall reasons.\n"
             f"* `{out}_ncu_dram_steady.json` — 16 consecutive launches in one pass with `--cache-control none`: DRAM bytes per launch of the back-to-back regime (what `roofline.traffic` quotes).\n"
             f"* `{out}_ncu_variants.json` — the same `--set full` numbers for VerletSGLD + fused prior and the all-sums variant (both TMA-staged), HMC, `initial_step(save_state)`.\n"
             f"* `{out}_ncu_launches.json` — launch list of `bench.py --steps 3 --warmup 3 --no-cpu --no-extra` (first 600 launches).\n"
             f"* `{out}_ncu_foreign_grads.json` — launch list of the reference runner's loop with gradients in tensors of their own: the step kernel only, no copy kernel.\n")
    add = os.path.join(P, f"{out}_addenda.md")
    if os.path.exists(add):
        L.append(open(add).read())
    open(os.path.join(P, f"{out}_notes.md"), "w").write("\n".join(L) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True, help="tag of the single-GPU session in gpurun_out/")
    ap.add_argument("--multi", help="tag of the multi-GPU session in gpurun_out/")
    ap.add_argument("--out", default="r02")
    a = ap.parse_args()
    t, o = a.tag, a.out
    note = ("final code of the round; step kernel of tools/ncu_target.py sgld, caches flushed between replays = production regime; "
            "launch list of bench.py --steps 3 --warmup 3 --no-cpu --no-extra")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "--tag", o, "--rep", os.path.join(G, f"{t}_step_cold.ncu-rep"),
                    "--launches", os.path.join(G, f"{t}_launches.csv"), "--dram-csv", os.path.join(G, f"{t}_dram_steady.csv"), "--note", note],
                   check=True, stdout=subprocess.DEVNULL)
    for src, dst in ((f"{t}_bench.json", f"{o}_bench.json"), (f"{t}_bench_20steps.json", f"{o}_bench_20steps.json"),
                     (f"{t}_bench_reference_arm.json", f"{o}_bench_reference_arm.json"), (f"{t}_small_models.json", f"{o}_small_models.json"),
                     (f"{t}_tune.json", f"{o}_tune.json"), (f"{t}_variants.jsonl", f"{o}_variants_final.jsonl")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    if a.multi:
        for n in (1, 2, 4, 8):
            src = os.path.join(G, f"{a.multi}_bench_n{n}_s20.json")
            if os.path.exists(src):
                shutil.copy(src, os.path.join(P, f"{o}_bench_n{n}.json"))
        src = os.path.join(G, f"{a.multi}_bench_n8_s200.json")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, f"{o}_bench_n8_200steps.json"))
    ncu_variants(t, o)
    parity_md(t, o)
    notes_md(o, bool(a.multi))
    print(open(os.path.join(P, f"{o}_notes.md")).read()[:3000])


if __name__ == "__main__":
    main()
