#!/usr/bin/env python
"""bench.py - A-step candidate pairs/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the Hi-C A-step hot path (get_actdist for every
candidate pair, igm/steps/ActivationDistanceStep.py:215-219,336-485) over the
candidate list of the final sigma (0.01) of the sweep on a synthetic population:
config 2 of BASELINE.json - 1 000 structures x 200 kb male diploid (29 838
beads).  For N > 1 every rank holds the full population (replicated), works on
its own equal-count shard of candidate pairs and the per-pair results are
collected with one NCCL all-gather (weak scaling: pairs per GPU fixed).

Prints ONE JSON line (rank 0).  `value` = whole-job pairs/s with inputs
resident in HBM; `e2e` = the same through the C-ABI host entry point
(igmk_actdist_host) with pinned host buffers, H2D/D2H inside the timed region;
`roofline` = algorithmic bytes / kernel time against the measured HBM copy
bandwidth; `cpu_baseline` = the oracle port of the reference's get_actdist on
the host cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "A-step candidate pairs/sec"
UNIT = "pairs/s"
SEED = 20261018


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nstruct", type=int, default=1000)
    ap.add_argument("--resolution", type=int, default=200_000)
    ap.add_argument("--sigma", type=float, default=0.01)
    ap.add_argument("--mode", default="LB", choices=["LB", "GP"])
    ap.add_argument("--it-corr", type=int, default=0)
    ap.add_argument("--max-pairs", type=int, default=0, help="truncate the candidate list (debug)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ workload
def build_index_and_pairs(args, rank):
    """Candidate pairs of this rank: the sigma-filtered non-zeros of the synthetic
    probability matrix in CSR order (what setup() produces,
    ActivationDistanceStep.py:171-178); ranks > 0 take the same list rotated
    along the genome so every GPU does different, equally sized work."""
    from igm_b200 import synthetic
    from igm_b200.steps.ActivationDistanceStep import filter_candidates
    bins = synthetic.genome_bins(args.resolution)
    chrom_hap, chrom_bead, copy_bead, ci = synthetic.build_index(bins)
    pm = synthetic.make_prob_matrix(chrom_hap, seed=SEED)
    ii, jj, pw = filter_candidates(pm, args.sigma, args.sigma)
    if args.max_pairs:
        ii, jj, pw = ii[:args.max_pairs], jj[:args.max_pairs], pw[:args.max_pairs]
    if rank:
        n = len(chrom_hap)
        sh = (rank * 977) % n
        a, b = (ii.astype(np.int64) + sh) % n, (jj.astype(np.int64) + sh) % n
        # keep the reference's precondition for LB intra pairs: equal copy counts
        nc = ci.ncopies()
        ok = ~((chrom_hap[a] == chrom_hap[b]) & (nc[a] != nc[b])) & (a != b)
        a[~ok], b[~ok] = ii[~ok], jj[~ok]
        ii, jj = a.astype(np.int32), b.astype(np.int32)
    return chrom_hap, chrom_bead, copy_bead, ci, ii, jj, pw


def config_name(args):
    """Which BASELINE.json config the arguments correspond to."""
    if args.resolution == 200_000 and args.nstruct == 1000:
        return "config 2"
    if args.resolution == 200_000 and args.nstruct == 10000:
        return "config 3 (per-GPU shard)"
    if args.resolution == 50_000 and args.nstruct == 1000:
        return "config 5"
    return "custom"


def algorithmic_bytes(ci, ii, jj, nstruct):
    """B_pair = 12 N (c_i + c_j) + 24 (input record) + 24 (result)."""
    nc = ci.ncopies().astype(np.int64)
    return int((12 * nstruct * (nc[ii] + nc[jj]) + 48).sum())


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (one long-lived
    `nvidia-smi -lms 50` process; B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None
        self.t_start = None
        self.t_stop = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def mark_start(self):
        self.t_start = time.perf_counter()

    def mark_stop(self):
        self.t_stop = time.perf_counter()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=5)

    def summary(self):
        rows = [r for t, r in self.rows
                if (self.t_start is None or t >= self.t_start) and (self.t_stop is None or t <= self.t_stop + 0.06)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names)
                   if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def measured_traffic(nstruct, n_pairs, mode, key="dram_bytes_per_launch"):
    """DRAM bytes (read + write) of ONE launch of the dominant kernel on this
    workload, from the committed `ncu --set full` capture (profiles/traffic.json,
    written by profiles/update_traffic.py); None when no capture matches.  ``key`` selects
    another per-launch counter of the same capture (warp instructions)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        for e in json.load(open(p))["captures"]:
            if e["nstruct"] == nstruct and e["n_pairs"] == n_pairs and e["mode"] == mode:
                return float(e[key])
    except Exception:
        pass
    return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------- CPU baseline
_G = {}


def _cpu_worker(sl):
    from oracle import actdist_oracle as orc
    ii, jj, pw, pl, coords, radii, chrom_hap, ci, it_corr, mode = _G["a"]
    a, b = sl
    recs, _ = orc.run_pairs(ii[a:b], jj[a:b], pw[a:b], pl[a:b], coords, radii, chrom_hap, ci,
                            it_corr, 2.0, mode)
    return len(recs)


def cpu_baseline(sample_coords, radii, chrom_hap, ci, ii, jj, pw, it_corr, mode, budget_s):
    """The oracle's NumPy port of the reference get_actdist (same NumPy calls as
    ActivationDistanceStep.py:405-473), batches of 1000 pairs as in setup (:129),
    multiprocessing over all host cores, coordinates shared by fork."""
    import multiprocessing as mp
    from oracle import actdist_oracle as orc
    cores = os.cpu_count() or 1
    pl = np.zeros(len(ii))
    omode = orc.MODE_LB if mode == "LB" else orc.MODE_GP
    _G["a"] = (ii, jj, pw, pl, sample_coords, radii, chrom_hap, ci, it_corr, omode)
    t0 = time.perf_counter()
    _cpu_worker((0, min(200, len(ii))))
    rate1 = min(200, len(ii)) / (time.perf_counter() - t0)
    n = int(min(len(ii), max(1000, rate1 * cores * budget_s * 0.8)))
    batches = [(a, min(n, a + 1000)) for a in range(0, n, 1000)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, batches, chunksize=1)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "first %d candidate pairs of the step's list, NumPy oracle port of get_actdist, "
                      "fork pool of %d processes, batches of 1000; %.1f s" % (n, cores, dt),
            "single_core_pairs_per_s": rate1}


def sample_population_host(eng_coords_t, beads):
    return eng_coords_t[beads].cpu().numpy()


# ------------------------------------------------------------------- main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    from igm_b200 import synthetic, _lib
    from igm_b200.engine import ActdistEngine, launch_count

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; igm_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    chrom_hap, chrom_bead, copy_bead, ci, ii, jj, pw = build_index_and_pairs(args, rank)
    nbead = len(chrom_bead)
    n_pairs = len(ii)
    radius = float(synthetic.bead_radius(nbead))
    radii = np.full(nbead, radius, np.float32)

    # population: generated on the device, then staged into the engine's layout
    coords_t = synthetic.random_walk_coordinates_torch(chrom_bead, copy_bead, args.nstruct, radius,
                                                       SEED, dev)
    eng = ActdistEngine(nbead=nbead, nstruct=args.nstruct, device=local_rank)
    eng.upload_coordinates(coords_t)
    eng.set_index(ci.ptr, ci.beads, chrom_hap, radii)

    # device-resident inputs; results land directly in this rank's slice of the
    # all-gather buffer (in-place NCCL all-gather)
    d_i = torch.from_numpy(ii).to(dev)
    d_j = torch.from_numpy(jj).to(dev)
    d_pw = torch.from_numpy(pw).to(dev)
    d_pl = torch.zeros(n_pairs, dtype=torch.float64, device=dev)
    gather = torch.zeros((world, n_pairs, 32), dtype=torch.uint8, device=dev)
    mine = gather[rank]
    stream = torch.cuda.current_stream().cuda_stream

    # N > 1: the kernel stores its results straight into every GPU's gather buffer
    # over NVLink (peer-mapped memory); NCCL all-gather is the fallback
    from igm_b200.dist import PeerGather, peer_gather_available
    pg = None
    if world > 1 and peer_gather_available(world):
        try:
            pg = PeerGather(n_pairs, rank, world, dev)
        except Exception as e:                       # no peer mapping on this box: NCCL all-gather
            if rank == 0:
                sys.stderr.write("peer gather unavailable (%s); using NCCL all-gather\n" % e)
            pg = None
        ok = torch.tensor([1 if pg is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)    # all ranks must agree on the path
        if not int(ok.item()):
            pg = None

    def step():
        if pg is not None:
            return pg.step(eng, d_i, d_j, d_pw, d_pl, n_pairs, 2.0, args.it_corr, args.mode, stream)
        eng.actdist_device(d_i, d_j, d_pw, d_pl, mine, n_pairs, 2.0, args.it_corr, args.mode,
                           stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(gather.view(world * n_pairs, 32), mine)
        return gather

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark_start()
    l0 = launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    ev[0].record()
    last = gather
    for k in range(args.steps):
        ev[2 * k + 1].record()
        if pg is not None:
            b = pg.k % len(pg.bufs)
            pg.k += 1
            eng.actdist_device_peers(d_i, d_j, d_pw, d_pl, pg.peer_slices[b], world, n_pairs, 2.0,
                                     args.it_corr, args.mode, stream)
            ev[2 * k + 2].record()
            pg.hdls[b].barrier()
            eng.finish_results(pg.bufs[b], world * n_pairs, stream)
            last = pg.bufs[b]
        else:
            eng.actdist_device(d_i, d_j, d_pw, d_pl, mine, n_pairs, 2.0, args.it_corr, args.mode,
                               stream=stream)
            ev[2 * k + 2].record()
            if world > 1:
                dist.all_gather_into_tensor(gather.view(world * n_pairs, 32), mine)
    ev[-1].record()
    torch.cuda.synchronize()
    sampler.mark_stop()          # moved further out below when the e2e region runs too
    if world > 1:
        dist.barrier()
    launches = launch_count() - l0
    redo_pairs = eng.last_redo_count()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n_pairs / (ms_per_step * 1e-3)

    # ---- the other A-steps of the sigma sweep of configs[1] (demo/config_file.json:35-50),
    # outside the timed region: one candidate list per sigma, device-resident, 3 launches each
    sweep = None
    if rank == 0 and world == 1 and not args.max_pairs and abs(args.sigma - 0.01) < 1e-12:
        from igm_b200.steps.ActivationDistanceStep import filter_candidates as _fc
        pm_s = synthetic.make_prob_matrix(chrom_hap, seed=SEED)
        sweep = {"sigmas": [], "note": "one A-step per sigma; device-resident, mean of 3 launches after 1 warm-up"}
        tot_pairs, tot_ms = n_pairs, ms_per_step
        for sg in (1.0, 0.2, 0.1, 0.05, 0.02):
            si, sj, sp = _fc(pm_s, sg, sg)
            t_i, t_j = torch.from_numpy(si).to(dev), torch.from_numpy(sj).to(dev)
            t_p = torch.from_numpy(sp).to(dev)
            t_l = torch.zeros(len(si), dtype=torch.float64, device=dev)
            t_o = torch.zeros((len(si), 32), dtype=torch.uint8, device=dev)
            eng.actdist_device(t_i, t_j, t_p, t_l, t_o, len(si), 2.0, args.it_corr, args.mode, stream=stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                eng.actdist_device(t_i, t_j, t_p, t_l, t_o, len(si), 2.0, args.it_corr, args.mode, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3.0
            sweep["sigmas"].append({"sigma": sg, "pairs": int(len(si)), "ms": ms,
                                    "pairs_per_s": len(si) / (ms * 1e-3)})
            tot_pairs += len(si)
            tot_ms += ms
        sweep["sigmas"].append({"sigma": 0.01, "pairs": int(n_pairs), "ms": ms_per_step, "pairs_per_s": value})
        sweep["sweep_pairs"] = int(tot_pairs)
        sweep["sweep_ms"] = tot_ms
        sweep["sweep_pairs_per_s"] = tot_pairs / (tot_ms * 1e-3)
        del pm_s

    # ---- parity gate inside the bench: a sample of pairs against the oracle
    parity = None
    if rank == 0:
        from oracle import actdist_oracle as orc
        rng = np.random.default_rng(1)
        sel = np.sort(rng.choice(n_pairs, size=min(300, n_pairs), replace=False))
        got = last[rank].cpu().numpy().reshape(-1).view(_lib.PAIR_RESULT_DTYPE)[sel]
        hap_needed = np.unique(np.concatenate([ii[sel], jj[sel]]))
        beads = np.unique(np.concatenate([ci[h] for h in hap_needed]))
        remap = -np.ones(nbead, np.int64)
        remap[beads] = np.arange(len(beads))
        sub_coords = coords_t[torch.from_numpy(beads).to(dev)].cpu().numpy()

        class _CI:
            def __getitem__(self, i):
                return [int(remap[b]) for b in ci[i]]
        _, dets = orc.run_pairs(ii[sel], jj[sel], pw[sel], np.zeros(len(sel)), sub_coords,
                                radii[beads], chrom_hap, _CI(), args.it_corr, 2.0,
                                orc.MODE_LB if args.mode == "LB" else orc.MODE_GP)
        exp = orc.details_to_arrays(dets)
        parity = bool(np.array_equal(got["d2_sel_bits"], exp["d2_sel_bits"])
                      and np.array_equal(got["contact_count"], exp["contact_count"])
                      and np.array_equal(got["o"], exp["o"])
                      and np.array_equal(got["p"].view(np.uint64), exp["p"].view(np.uint64)))

    # ---- e2e: C-ABI host entry point, pinned host buffers, copies inside
    e2e = None
    if not args.no_e2e:
        h_i = torch.from_numpy(ii).pin_memory()
        h_j = torch.from_numpy(jj).pin_memory()
        h_pw = torch.from_numpy(pw).pin_memory()
        h_pl = torch.zeros(n_pairs, dtype=torch.float64).pin_memory()
        h_out = torch.zeros(n_pairs * 32, dtype=torch.uint8).pin_memory()
        lib = _lib.load()

        def e2e_step():
            _lib.check(lib.igmk_actdist_host(eng._ctx, n_pairs, h_i.data_ptr(), h_j.data_ptr(),
                                             h_pw.data_ptr(), h_pl.data_ptr(), 2.0, args.it_corr,
                                             0 if args.mode == "LB" else 1, 0, h_out.data_ptr()))
        for _ in range(2):
            e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        dt = time.perf_counter() - t0
        sampler.mark_stop()      # the clock samples cover both timed regions
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": world * n_pairs * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(n_pairs * 24), "d2h_bytes_per_step": int(n_pairs * 32),
               "api": "igmk_actdist_host (C ABI) with pinned host buffers"}

    if rank == 0:
        time.sleep(0.1)
        sampler.stop()
        peak, peak_src = measured_peak_gbs()
        alg = algorithmic_bytes(ci, ii, jj, args.nstruct)
        k_ms = float(np.mean(kernel_ms))
        achieved = alg / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "%s: synthetic %d-structure population at %d kb male diploid "
                            "(%d beads), Hi-C A-step (%s) over the sigma=%g candidate list, "
                            "%d pairs per GPU" % (config_name(args), args.nstruct, args.resolution // 1000,
                                                  nbead, args.mode, args.sigma, n_pairs),
                "pairs_per_gpu": n_pairs, "nstruct": args.nstruct, "nbead": nbead,
                "pair_structs_per_s": value * args.nstruct,
                "parallelism": "pairs sharded over %d GPU(s), coordinates replicated, %s" % (
                    world, "results stored into every GPU's gather buffer from inside the kernel "
                    "(NVLink peer stores) + one barrier" if pg is not None else
                    "one NCCL all-gather of results"),
                "l2": "inputs larger than L2 (coordinates %.0f MB, pair list %.0f MB)" % (
                    nbead * 3 * eng.nstruct * 4 / 1e6, n_pairs * 24 / 1e6),
                "parity_sample_ok": parity,
                "list_form_redo_pairs": redo_pairs,
                "sigma_sweep": sweep,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic(args.nstruct, n_pairs, args.mode),
                         "kernel": "actdist_warp_kernel" if args.nstruct <= 1024 else "actdist_block_kernel",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg,
                         "peak_source": peak_src},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        # what actually limits the kernel: the issue slots it uses,
        # from the warp-instruction count of the committed ncu capture and the live kernel time
        winst = measured_traffic(args.nstruct, n_pairs, args.mode, "warp_instructions_per_launch")
        sm_mhz = line["clocks"].get("sm_mhz") or 1965.0
        if winst:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            line["roofline"]["issue"] = {
                "warp_instructions_per_launch": winst,
                "achieved_winst_per_s": winst / (k_ms * 1e-3),
                "peak_winst_per_s": sms * 4 * sm_mhz * 1e6,
                "frac": winst / (k_ms * 1e-3) / (sms * 4 * sm_mhz * 1e6),
                "note": "issue slots used (4 schedulers per SM at the sampled SM clock); the kernel is latency-bound"}
        # ... and the unit it keeps busiest: the L1 / shared-memory data pipe (LSU wavefronts,
        # 128 bytes per clock per SM) - ncu figure of the committed capture, rescaled to the
        # live kernel time
        l1pct = measured_traffic(args.nstruct, n_pairs, args.mode, "l1tex_lsu_data_pipe_pct")
        ms_ncu = measured_traffic(args.nstruct, n_pairs, args.mode, "kernel_ms_under_ncu")
        if l1pct and ms_ncu:
            line["roofline"]["l1tex"] = {
                "lsu_data_pipe_frac": l1pct / 100.0 * ms_ncu / k_ms,
                "lsu_data_pipe_frac_under_ncu": l1pct / 100.0, "kernel_ms_under_ncu": ms_ncu,
                "note": "l1tex__data_pipe_lsu_wavefronts: coordinate loads, tile and key-array traffic"}
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
            # bounded CPU sample (about --cpu-seconds of work on all host cores):
            # needs host copies of the beads the sample touches
            nb = 320000
            hap_needed = np.unique(np.concatenate([ii[:nb * 8], jj[:nb * 8]]))
            beads = np.unique(np.concatenate([ci[h] for h in hap_needed]))
            remap = -np.ones(nbead, np.int64)
            remap[beads] = np.arange(len(beads))
            sub = coords_t[torch.from_numpy(beads).to(dev)].cpu().numpy()

            class _CI2:
                def __getitem__(self, i):
                    return [int(remap[b]) for b in ci[i]]
            line["cpu_baseline"] = cpu_baseline(sub, radii[beads], chrom_hap, _CI2(), ii[:nb * 8],
                                                jj[:nb * 8], pw[:nb * 8], args.it_corr, args.mode,
                                                args.cpu_seconds)
            # second CPU figure and a much larger parity gate: the plain-C oracle (OpenMP,
            # all host threads) over the same sample, compared with the GPU results of the
            # last timed step pair by pair
            from oracle import c_oracle
            if c_oracle.available():
                n_c = int(min(len(ii), nb * 8, 1_000_000))
                t0 = time.perf_counter()
                exp_c = c_oracle.run_pairs(ii[:n_c], jj[:n_c], pw[:n_c], np.zeros(n_c), sub, radii[beads],
                                           chrom_hap, ci.ptr, remap[ci.beads].astype(np.int32),
                                           args.it_corr, 2.0, 0 if args.mode == "LB" else 1)
                dt_c = time.perf_counter() - t0
                got_c = last[rank][:n_c].cpu().numpy().reshape(-1).view(_lib.PAIR_RESULT_DTYPE)
                ok = exp_c["o"] >= 0
                equal = bool(np.array_equal(got_c["d2_sel_bits"][ok], exp_c["d2_sel_bits"][ok])
                             and np.array_equal(got_c["contact_count"], exp_c["contact_count"])
                             and np.array_equal(got_c["o"], exp_c["o"])
                             and np.array_equal(got_c["p"].view(np.uint64), exp_c["p"].view(np.uint64))
                             and np.array_equal(got_c["nrec"], exp_c["nrec"]))
                line["cpu_baseline"]["c_port"] = {
                    "value": n_c / dt_c, "unit": UNIT, "cores": os.cpu_count() or 1,
                    "sample": "first %d pairs, oracle/actdist_oracle.c with OpenMP; %.1f s" % (n_c, dt_c),
                    "gpu_results_equal": equal}
                line["config"]["parity_c_oracle_pairs"] = n_c
                line["config"]["parity_c_oracle_ok"] = equal
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle
    port of get_actdist; the reference itself is Python and needs alabtools/h5py
    to run end to end) on the host cores, same config/metric; bounded sample per
    step.  Under torchrun only rank 0 works."""
    if rank != 0:
        return
    from igm_b200 import synthetic
    chrom_hap, chrom_bead, copy_bead, ci, ii, jj, pw = build_index_and_pairs(args, 0)
    nbead = len(chrom_bead)
    radius = float(synthetic.bead_radius(nbead))
    cores = os.cpu_count() or 1
    # bounded sample: pairs of the first rows of the list; only their beads are generated
    n_sample = int(min(len(ii), 25000 * cores))
    si, sj, spw = ii[:n_sample], jj[:n_sample], pw[:n_sample]
    hap_needed = np.unique(np.concatenate([si, sj]))
    beads = np.unique(np.concatenate([ci[h] for h in hap_needed]))
    remap = -np.ones(nbead, np.int64)
    remap[beads] = np.arange(len(beads))
    rng = np.random.default_rng(SEED)
    sub = synthetic.random_walk_coordinates(chrom_bead[beads], copy_bead[beads], args.nstruct,
                                            radius, rng)

    class _CI:
        def __getitem__(self, i):
            return [int(remap[b]) for b in ci[i]]
    radii = np.full(len(beads), radius, np.float32)
    per_step, per_ms = [], []
    info = None
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        info = cpu_baseline(sub, radii, chrom_hap, _CI(), si, sj, spw, args.it_corr, args.mode,
                            max(2.0, args.cpu_seconds / max(1, args.steps)))
        if k >= args.warmup:
            per_step.append(info["value"])
            per_ms.append((time.perf_counter() - t0) * 1e3)
    value = float(np.mean(per_step))
    # for context: the plain-C oracle (OpenMP) on the same sample
    c_port = None
    try:
        from oracle import c_oracle
        if c_oracle.available():
            n_c = int(min(len(si), 400000))
            t0 = time.perf_counter()
            c_oracle.run_pairs(si[:n_c], sj[:n_c], spw[:n_c], np.zeros(n_c), sub, radii, chrom_hap, ci.ptr,
                               remap[ci.beads].astype(np.int32), args.it_corr, 2.0,
                               0 if args.mode == "LB" else 1)
            c_port = {"value": n_c / (time.perf_counter() - t0), "unit": UNIT, "cores": cores,
                      "sample": "first %d pairs, oracle/actdist_oracle.c with OpenMP" % n_c}
    except Exception as e:          # the C oracle is optional here
        c_port = {"unavailable": str(e)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(per_ms)),
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: synthetic %d-structure population at %d kb male diploid "
                               "(%d beads), Hi-C A-step (%s) over the sigma=%g candidate list; "
                               "bounded CPU sample per step" % (config_name(args), args.nstruct,
                                                                args.resolution // 1000,
                                                                nbead, args.mode, args.sigma),
                   "nstruct": args.nstruct, "nbead": nbead},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "port",
                         "sample": info["sample"], "c_port": c_port},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
