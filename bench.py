#!/usr/bin/env python
"""bench.py - A-step candidate pairs/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the Hi-C A-step hot path (get_actdist for every
candidate pair, igm/steps/ActivationDistanceStep.py:215-219,336-485) over the
candidate list of the final sigma (0.01) of the sweep on a synthetic population:
config 2 of BASELINE.json - 1 000 structures x 200 kb male diploid (29 838
beads).  The population comes from the seeded NumPy generator, so both arms of the
bench (ours / --impl reference) work on the SAME coordinates and the SAME list.
For N > 1 every rank holds the full population (replicated), works on its own
equal-count list (weak scaling) and the per-pair results are gathered on every
GPU (peer stores from inside the kernel, or one NCCL all-gather).

Prints ONE JSON line (rank 0).  `value` = whole-job pairs/s with inputs resident
in HBM; `e2e` = the same through the C ABI with HOST buffers: the population is
staged again every step (every A-step follows an M-step that rewrote the .hss)
by igmk_actdist_host_population, which pipelines the upload of the pinned host
coordinates with the pair kernels, copies the pair list in and the results out;
`e2e_two_calls` = igmk_upload_coords followed by igmk_actdist_host;
`e2e_pairs_only` leaves the population resident; `roofline` = algorithmic bytes / kernel time against the measured HBM
copy bandwidth; `cpu_baseline` = the reference's own get_actdist on the host
cores (bounded sample).  `config.extra` carries the other BASELINE.json configs,
measured in the same run: config 3 (10 000 structures, one fixed list strong-scaled
over the GPUs), config 4 (contact-frequency map) and config 5 (50 kb stress test).
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
N_ALL_HAPLOID_PAIRS_200KB = 15453 * 15452 // 2       # "all candidate pairs" upper bound of config 3


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
    ap.add_argument("--no-extras", action="store_true", help="skip configs 3 / 4 / 5")
    ap.add_argument("--extras", default="3,4,5", help="which extra configs to run")
    return ap.parse_args()


# ------------------------------------------------------------------ workload
def build_index_and_pairs(resolution, sigma, rank=0, max_pairs=0, native=True):
    """Candidate pairs of this rank: the sigma-filtered non-zeros of the synthetic
    probability matrix in CSR order (what setup() produces,
    ActivationDistanceStep.py:171-178); ranks > 0 take the same list rotated
    along the genome so every GPU does different, equally sized work."""
    from igm_b200 import synthetic
    from igm_b200.steps.ActivationDistanceStep import filter_candidates
    bins = synthetic.genome_bins(resolution)
    chrom_hap, chrom_bead, copy_bead, ci = synthetic.build_index(bins)
    pm = synthetic.make_prob_matrix(chrom_hap, seed=SEED)
    ii, jj, pw = filter_candidates(pm, sigma, sigma, native=native)
    if max_pairs:
        ii, jj, pw = ii[:max_pairs], jj[:max_pairs], pw[:max_pairs]
    if rank:
        n = len(chrom_hap)
        sh = (rank * 977) % n
        a, b = (ii.astype(np.int64) + sh) % n, (jj.astype(np.int64) + sh) % n
        # keep the reference's precondition for LB intra pairs: equal copy counts
        nc = ci.ncopies()
        ok = ~((chrom_hap[a] == chrom_hap[b]) & (nc[a] != nc[b])) & (a != b)
        a[~ok], b[~ok] = ii[~ok], jj[~ok]
        ii, jj = a.astype(np.int32), b.astype(np.int32)
    return chrom_hap, chrom_bead, copy_bead, ci, ii, jj, pw, pm


def host_population(chrom_bead, copy_bead, nstruct, radius):
    """(nbead, nstruct, 3) float32 from the seeded NumPy generator - identical in both
    arms of the bench."""
    from igm_b200 import synthetic
    return synthetic.random_walk_coordinates(chrom_bead, copy_bead, nstruct, radius, np.random.default_rng(SEED))


def config_name(resolution, nstruct):
    if resolution == 200_000 and nstruct == 1000:
        return "config 2"
    if resolution == 200_000 and nstruct == 10000:
        return "config 3"
    if resolution == 50_000 and nstruct == 1000:
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
    """Counter of ONE launch of the dominant kernel on this workload from the committed
    `ncu --set full` capture (profiles/traffic.json, written by profiles/update_traffic.py);
    None when no capture matches."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        for e in json.load(open(p))["captures"]:
            if e["nstruct"] == nstruct and e["n_pairs"] == n_pairs and e["mode"] == mode and key in e:
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


# ------------------------------------------------------------- CPU arms
_G = {}


def _ref_worker(sl):
    """One batch of the reference's task() loop (ActivationDistanceStep.py:215-222) with
    the reference's own, unmodified get_actdist."""
    fn, hss, ii, jj, pw, it_corr = _G["a"]
    a, b = sl
    n = 0
    for k in range(a, b):
        n += len(fn(int(ii[k]), int(jj[k]), pw[k], 0.0, hss, it_corr, contactRange=2.0))
    return n


def cpu_reference_rate(coords, radii, chrom_hap, ci, ii, jj, pw, it_corr, mode, budget_s, steps=1, warmup=0):
    """The reference's own get_actdist (imported unmodified through oracle/ref_loader.py)
    on all host cores: batches of 1000 pairs as in setup (:129) over a fork pool that is
    created BEFORE the timed span; coordinates inherited by fork.  The sample is the
    leading slice of the step's list, sized to about `budget_s` seconds per step."""
    import multiprocessing as mp
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    fn = ref["lb_get_actdist"] if mode == "LB" else ref["gp_get_actdist"]
    hss = ref_loader.FakeHss(coords, radii, chrom_hap, ci)
    cores = os.cpu_count() or 1
    root = ref_loader.REFERENCE_ROOT
    nmax = min(len(ii), 4_000_000)
    pw64 = np.ascontiguousarray(pw[:nmax], np.float64)      # pw64[k] is the np.float64 setup() hands over
    _G["a"] = (fn, hss, ii, jj, pw64, it_corr)
    t0 = time.perf_counter()
    _ref_worker((0, min(100, nmax)))
    rate1 = min(100, nmax) / (time.perf_counter() - t0)
    n = int(min(nmax, len(pw64), max(1000, rate1 * cores * budget_s * 0.6)))
    batches = [(a, min(n, a + 1000)) for a in range(0, n, 1000)]
    rates, ms = [], []
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_ref_worker, [(0, 1)] * cores, chunksize=1)          # the workers are up
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, batches, chunksize=1)
            dt = time.perf_counter() - t0
            if k >= warmup:
                rates.append(n / dt)
                ms.append(dt * 1e3)
    where = os.path.relpath(root, ROOT) if root.startswith(ROOT) else root
    return {"value": float(np.mean(rates)), "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "first %d candidate pairs of the step's list through the reference's unmodified "
                      "get_actdist (%s/igm/steps/ActivationDistanceStep.py), fork pool of %d processes created "
                      "before the timed span, batches of 1000; %.1f s per step" % (
                          n, where, cores, float(np.mean(ms)) / 1e3),
            "single_core_pairs_per_s": rate1, "sample_pairs": n, "ms_per_step": float(np.mean(ms))}


def c_oracle_check(ii, jj, pw, coords_sub, radii_sub, chrom_hap, ci, remap, it_corr, mode, got):
    """Plain-C oracle (OpenMP) over the given pairs, compared field by field with `got`."""
    from oracle import c_oracle
    t0 = time.perf_counter()
    exp = c_oracle.run_pairs(ii, jj, pw, np.zeros(len(ii)), coords_sub, radii_sub, chrom_hap, ci.ptr,
                             remap[ci.beads].astype(np.int32), it_corr, 2.0, 0 if mode == "LB" else 1)
    dt = time.perf_counter() - t0
    ok = exp["o"] >= 0
    equal = bool(np.array_equal(got["d2_sel_bits"][ok], exp["d2_sel_bits"][ok])
                 and np.array_equal(got["contact_count"], exp["contact_count"])
                 and np.array_equal(got["o"], exp["o"])
                 and np.array_equal(got["p"].view(np.uint64), exp["p"].view(np.uint64))
                 and np.array_equal(got["nrec"], exp["nrec"]))
    return equal, dt


def sub_population(coords_get, ci, nbead, haps):
    """Host copies of the beads the loci `haps` touch + the bead remap."""
    haps = np.unique(haps)
    nc = ci.ptr[haps + 1] - ci.ptr[haps]
    beads = np.unique(np.concatenate([ci.beads[ci.ptr[haps]], ci.beads[ci.ptr[haps[nc > 1]] + 1]]))
    remap = -np.ones(nbead, np.int64)
    remap[beads] = np.arange(len(beads))
    return beads, remap, coords_get(beads)


# ------------------------------------------------------------------- extras
def extra_config3(args, rank, world, dev, local_rank, dist, torch):
    """Config 3: 10 000 structures x 200 kb; ONE fixed list (every stored non-zero of the
    synthetic matrix = all candidate pairs) split into contiguous 1/N shares (strong
    scaling); results of all ranks gathered on every GPU."""
    from igm_b200 import synthetic, _lib
    from igm_b200.engine import ActdistEngine, launch_count
    from igm_b200.dist import shard_bounds
    nstruct = 10000
    chrom_hap, chrom_bead, copy_bead, ci, ii, jj, pw, _ = build_index_and_pairs(200_000, 0.0, 0, args.max_pairs)
    nbead, n_all = len(chrom_bead), len(ii)
    radius = float(synthetic.bead_radius(nbead))
    radii = np.full(nbead, radius, np.float32)
    coords_t = synthetic.random_walk_coordinates_torch(chrom_bead, copy_bead, nstruct, radius, SEED + 3, dev)
    eng = ActdistEngine(nbead=nbead, nstruct=nstruct, device=local_rank)
    eng.upload_coordinates(coords_t)
    eng.set_index(ci.ptr, ci.beads, chrom_hap, radii)
    per, bounds = shard_bounds(n_all, world)
    lo, hi = bounds[rank]
    n_mine = hi - lo
    d_i = torch.from_numpy(ii[lo:hi]).to(dev)
    d_j = torch.from_numpy(jj[lo:hi]).to(dev)
    d_pw = torch.from_numpy(pw[lo:hi]).to(dev)
    d_pl = torch.zeros(n_mine, dtype=torch.float64, device=dev)
    gather = torch.zeros((world, per, 32), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        eng.actdist_device(d_i, d_j, d_pw, d_pl, gather[rank], n_mine, 2.0, args.it_corr, args.mode, stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(gather.view(world * per, 32), gather[rank])
    step()
    torch.cuda.synchronize()
    steps = 3
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 2)]
    ev[0].record()
    for k in range(steps):
        ev[2 * k + 1].record()
        eng.actdist_device(d_i, d_j, d_pw, d_pl, gather[rank], n_mine, 2.0, args.it_corr, args.mode, stream=stream)
        ev[2 * k + 2].record()
        if world > 1:
            dist.all_gather_into_tensor(gather.view(world * per, 32), gather[rank])
    ev[-1].record()
    torch.cuda.synchronize()
    launches = launch_count() - l0
    redo = eng.last_redo_count()
    t = torch.tensor([ev[0].elapsed_time(ev[-1])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    k_ms = float(np.mean([ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(steps)]))
    out = None
    if rank == 0:
        value = n_all / (ms_per_step * 1e-3)
        peak, _ = measured_peak_gbs()
        alg = algorithmic_bytes(ci, ii[lo:hi], jj[lo:hi], nstruct)
        # parity: a sample of EVERY rank's slice of rank 0's gathered buffer against the C oracle
        rng = np.random.default_rng(3)
        n_chk = 24000
        sel = np.sort(np.concatenate([blo + rng.choice(bhi - blo, size=min(n_chk // world, bhi - blo), replace=False)
                                      for blo, bhi in bounds if bhi > blo]))
        rows = np.concatenate([r * per + (sel[(sel >= blo) & (sel < bhi)] - blo) for r, (blo, bhi) in enumerate(bounds)])
        got = gather.view(world * per, 32)[torch.from_numpy(rows).to(dev)].cpu().numpy().reshape(-1).view(
            _lib.PAIR_RESULT_DTYPE)
        beads, remap, sub = sub_population(lambda b: coords_t[torch.from_numpy(b).to(dev)].cpu().numpy(), ci, nbead,
                                           np.concatenate([ii[sel], jj[sel]]))
        equal, dt_c = c_oracle_check(ii[sel], jj[sel], pw[sel], sub, radii[beads], chrom_hap, ci, remap,
                                     args.it_corr, args.mode, got)
        out = {"workload": "config 3: synthetic 10000-structure population at 200 kb male diploid (%d beads), Hi-C "
                           "A-step (%s) over ONE fixed list of %d pairs (every stored non-zero of the synthetic "
                           "matrix), contiguous 1/%d shares" % (nbead, args.mode, n_all, world),
               "value": value, "unit": UNIT, "scaling": "strong", "n_gpus": world, "steps": steps,
               "ms_per_step": ms_per_step, "pairs_total": n_all, "pairs_per_gpu": per,
               "kernel_ms_rank0": k_ms, "gpu_launches": int(launches), "list_form_redo_pairs_rank0": redo,
               "seconds_for_all_119M_haploid_pairs_at_this_rate": N_ALL_HAPLOID_PAIRS_200KB / value,
               "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": alg / (k_ms * 1e-3) / 1e9 / peak,
                            "traffic": measured_traffic(nstruct, n_mine, args.mode),
                            "algorithmic_bytes_per_launch": alg},
               "gather_parity_ok": equal, "parity_pairs_checked": int(len(sel)),
               "parity_note": "sample of every rank's slice of rank 0's gathered buffer vs the C oracle (%.1f s)" % dt_c}
    del coords_t, gather, d_i, d_j, d_pw, d_pl
    eng.close()
    torch.cuda.empty_cache()
    return out


def extra_config5(args, dev, local_rank, torch):
    """Config 5: 1 000 structures x 50 kb (119 k beads), sigma = 0.01 list; rank 0, one GPU."""
    from igm_b200 import synthetic, _lib
    from igm_b200.engine import ActdistEngine
    nstruct = 1000
    chrom_hap, chrom_bead, copy_bead, ci, ii, jj, pw, _ = build_index_and_pairs(50_000, 0.01, 0, args.max_pairs)
    nbead, n_pairs = len(chrom_bead), len(ii)
    radius = float(synthetic.bead_radius(nbead))
    radii = np.full(nbead, radius, np.float32)
    coords_t = synthetic.random_walk_coordinates_torch(chrom_bead, copy_bead, nstruct, radius, SEED + 5, dev)
    eng = ActdistEngine(nbead=nbead, nstruct=nstruct, device=local_rank)
    eng.upload_coordinates(coords_t)
    eng.set_index(ci.ptr, ci.beads, chrom_hap, radii)
    d_i, d_j, d_pw = (torch.from_numpy(x).to(dev) for x in (ii, jj, pw))
    d_pl = torch.zeros(n_pairs, dtype=torch.float64, device=dev)
    d_out = torch.zeros((n_pairs, 32), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    eng.actdist_device(d_i, d_j, d_pw, d_pl, d_out, n_pairs, 2.0, args.it_corr, args.mode, stream=stream)
    torch.cuda.synchronize()
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.actdist_device(d_i, d_j, d_pw, d_pl, d_out, n_pairs, 2.0, args.it_corr, args.mode, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    redo = eng.last_redo_count()
    rng = np.random.default_rng(5)
    sel = np.sort(rng.choice(n_pairs, size=min(20000, n_pairs), replace=False))
    got = d_out[torch.from_numpy(sel).to(dev)].cpu().numpy().reshape(-1).view(_lib.PAIR_RESULT_DTYPE)
    beads, remap, sub = sub_population(lambda b: coords_t[torch.from_numpy(b).to(dev)].cpu().numpy(), ci, nbead,
                                       np.concatenate([ii[sel], jj[sel]]))
    equal, _ = c_oracle_check(ii[sel], jj[sel], pw[sel], sub, radii[beads], chrom_hap, ci, remap, args.it_corr,
                              args.mode, got)
    peak, _ = measured_peak_gbs()
    alg = algorithmic_bytes(ci, ii, jj, nstruct)
    out = {"workload": "config 5: synthetic 1000-structure population at 50 kb male diploid (%d beads), Hi-C A-step "
                       "(%s) over the sigma=0.01 list, %d pairs, 1 GPU" % (nbead, args.mode, n_pairs),
           "value": n_pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "ms_per_step": ms,
           "pairs": n_pairs, "list_form_redo_pairs": redo,
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": measured_traffic(nstruct, n_pairs, args.mode)},
           "parity_c_oracle_ok": equal, "parity_pairs_checked": int(len(sel))}
    del coords_t, d_out
    eng.close()
    torch.cuda.empty_cache()
    return out


def extra_config4(eng, nbead, nstruct, rank, world, dev, dist, torch, coords_host, radius):
    """Config 4: population contact-frequency counts (K2, igm-report path) for all bead
    pairs of the config-2 population: block rows of the upper triangle dealt to the ranks
    (no collective: tiles are independent)."""
    from igm_b200.contact import row_blocks
    from igm_b200.engine import launch_count
    B = 448            # 67 block rows at 29 838 beads: the boustrophedon deal balances 8 ranks to ~1 %
    out = torch.zeros((B, nbead), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    mine = row_blocks(nbead, B, rank, world)
    eng.contact_counts_device(0, min(B, nbead), 0, min(B, nbead), out, 2.0, False, stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launch_count()
    e0.record()
    pairs = 0
    for _ in range(reps):
        pairs = 0
        for r0, r1 in mine:
            eng.contact_counts_device(r0, r1 - r0, r0, nbead - r0, out, 2.0, False, stream)
            pairs += (r1 - r0) * (nbead - r0)
    e1.record()
    torch.cuda.synchronize()
    launches = launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1) / reps, float(pairs)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms, pairs = float(tmax[0].item()), int(t[1].item())
    if rank != 0:
        return None
    ops = pairs * nstruct
    sms = torch.cuda.get_device_properties(dev).multi_processor_count * world
    peak = sms * 128 * 1.965e9
    res = {"workload": "config 4: contact-frequency counts (K2) of the config-2 population, all %d x %d bead pairs "
                       "(upper triangle at block granularity, %d block rows of %d) x %d structures" % (
                           nbead, nbead, (nbead + B - 1) // B, B, nstruct),
           "metric": "contact-frequency bead-pair-structs/s", "value": ops / (ms * 1e-3), "ms": ms, "n_gpus": world,
           "bead_pairs": pairs, "gpu_launches_rank0": int(launches), "reps": reps,
           "roofline": {"bound": "fp32-lanes", "achieved_laneops_per_s": 9 * ops / (ms * 1e-3),
                        "peak_laneops_per_s": peak, "frac": 9 * ops / (ms * 1e-3) / peak,
                        "note": "9 lane-operations per bead pair and structure (8 non-FMA float32 + 1 compare); "
                                "peak = SMs x 128 lanes x 1965 MHz"},
           "parity": "unpinned against alabtools (not in the reference tree); sample checked against "
                     "oracle/contact_oracle.py"}
    if coords_host is not None:
        from oracle import contact_oracle as co
        got = torch.zeros((64, 64), dtype=torch.int32, device=dev)
        eng.contact_counts_device(0, 64, 0, 64, got, 2.0, False, stream)
        torch.cuda.synchronize()
        exp = co.contact_counts_fast(np.ascontiguousarray(coords_host[:64]), np.full(64, radius, np.float32),
                                     np.arange(64), np.arange(64))
        res["parity_sample_ok"] = bool(np.array_equal(got.cpu().numpy().astype(np.uint32), exp))
    return res


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

    chrom_hap, chrom_bead, copy_bead, ci, ii, jj, pw, _ = build_index_and_pairs(
        args.resolution, args.sigma, rank, args.max_pairs)
    nbead = len(chrom_bead)
    n_pairs = len(ii)
    radius = float(synthetic.bead_radius(nbead))
    radii = np.full(nbead, radius, np.float32)

    # population: seeded NumPy generator (same coordinates as the reference arm) in pinned
    # host memory; larger populations (debug runs with --nstruct) are generated on the device
    host_pop = args.nstruct <= 2000
    coords_np = None
    if host_pop:
        coords_h = torch.from_numpy(host_population(chrom_bead, copy_bead, args.nstruct, radius)).pin_memory()
        coords_np = coords_h.numpy()

        def coords_get(beads):
            return coords_np[beads]
    else:
        coords_t = synthetic.random_walk_coordinates_torch(chrom_bead, copy_bead, args.nstruct, radius, SEED, dev)
        coords_h = None

        def coords_get(beads):
            return coords_t[torch.from_numpy(beads).to(dev)].cpu().numpy()
    eng = ActdistEngine(nbead=nbead, nstruct=args.nstruct, device=local_rank)
    eng.upload_coordinates(coords_h if host_pop else coords_t)
    eng.set_index(ci.ptr, ci.beads, chrom_hap, radii)

    # device-resident inputs; results land directly in this rank's slice of the gather buffer
    d_i = torch.from_numpy(ii).to(dev)
    d_j = torch.from_numpy(jj).to(dev)
    d_pw = torch.from_numpy(pw).to(dev)
    d_pl = torch.zeros(n_pairs, dtype=torch.float64, device=dev)
    gather = torch.zeros((world, n_pairs, 32), dtype=torch.uint8, device=dev)
    mine = gather[rank]
    stream = torch.cuda.current_stream().cuda_stream

    # N > 1: the kernel stores its (finished) records straight into every GPU's gather
    # buffer over NVLink (peer-mapped memory); NCCL all-gather is the fallback
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
        eng.actdist_device(d_i, d_j, d_pw, d_pl, mine, n_pairs, 2.0, args.it_corr, args.mode, stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(gather.view(world * n_pairs, 32), mine)
        return gather

    warm = max(args.warmup, 3)
    for _ in range(warm):
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
            last = pg.bufs[b]
        else:
            eng.actdist_device(d_i, d_j, d_pw, d_pl, mine, n_pairs, 2.0, args.it_corr, args.mode, stream=stream)
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

    # ---- parity gates inside the bench (rank 0): the NumPy oracle on a sample of the own slice,
    # and - N > 1 - the C oracle on a sample of EVERY rank's slice of rank 0's gathered buffer
    parity, gather_parity = None, None
    if rank == 0:
        from oracle import actdist_oracle as orc
        rng = np.random.default_rng(1)
        sel = np.sort(rng.choice(n_pairs, size=min(300, n_pairs), replace=False))
        got = last[rank].cpu().numpy().reshape(-1).view(_lib.PAIR_RESULT_DTYPE)[sel]
        beads, remap, sub_coords = sub_population(coords_get, ci, nbead, np.concatenate([ii[sel], jj[sel]]))

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
        if world > 1:
            from oracle import c_oracle
            if c_oracle.available():
                oks, checked = [], 0
                for r in range(world):
                    _, _, _, _, ri, rj, rp, _ = build_index_and_pairs(args.resolution, args.sigma, r, args.max_pairs)
                    s2 = np.sort(rng.choice(n_pairs, size=min(4000, n_pairs), replace=False))
                    g2 = last[r][torch.from_numpy(s2).to(dev)].cpu().numpy().reshape(-1).view(_lib.PAIR_RESULT_DTYPE)
                    b2, rm2, sc2 = sub_population(coords_get, ci, nbead, np.concatenate([ri[s2], rj[s2]]))
                    eq, _ = c_oracle_check(ri[s2], rj[s2], rp[s2], sc2, radii[b2], chrom_hap, ci, rm2, args.it_corr,
                                           args.mode, g2)
                    oks.append(eq)
                    checked += len(s2)
                gather_parity = {"ok": bool(all(oks)), "per_rank": oks, "pairs_checked": checked,
                                 "note": "sample of every rank's slice of rank 0's gathered buffer vs the C oracle"}

    # ---- e2e: C-ABI host entry points, pinned host buffers, copies inside the timed region
    e2e = e2e_pairs = None
    if not args.no_e2e:
        h_i = torch.from_numpy(ii).pin_memory()
        h_j = torch.from_numpy(jj).pin_memory()
        h_pw = torch.from_numpy(pw).pin_memory()
        h_pl = torch.zeros(n_pairs, dtype=torch.float64).pin_memory()
        h_out = torch.zeros(n_pairs * 32, dtype=torch.uint8).pin_memory()
        lib = _lib.load()

        def e2e_step(kind):
            a = (n_pairs, h_i.data_ptr(), h_j.data_ptr(), h_pw.data_ptr(), h_pl.data_ptr(), 2.0, args.it_corr,
                 0 if args.mode == "LB" else 1, 0, h_out.data_ptr())
            if kind == "pipelined":
                _lib.check(lib.igmk_actdist_host_population(eng._ctx, coords_h.data_ptr(), *a))
                return
            if kind == "replicated":
                # N > 1: each rank uploads 1 / N of the beads, one all-gather over NVLink
                from igm_b200.dist import replicate_population
                replicate_population(eng, coords_np_pinned, rank, world)
            if kind == "two_calls":
                _lib.check(lib.igmk_upload_coords(eng._ctx, coords_h.data_ptr(), 0))
            _lib.check(lib.igmk_actdist_host(eng._ctx, *a))

        def timed(kind):
            for _ in range(2):
                e2e_step(kind)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step(kind)
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        dt = timed("pairs_only")
        e2e_pairs = {"value": world * n_pairs * args.steps / dt, "unit": UNIT,
                     "h2d_bytes_per_step": int(n_pairs * 24), "d2h_bytes_per_step": int(n_pairs * 32),
                     "api": "igmk_actdist_host (C ABI), pinned host buffers, population resident in HBM"}
        e2e_two = None
        if host_pop:
            ref_out = None
            if rank == 0:
                e2e_step("two_calls")
                ref_out = h_out.clone()
            dt2 = timed("two_calls")
            dt = timed("pipelined")
            pop_mb = coords_h.numel() * 4 / 1e6
            e2e = {"value": world * n_pairs * args.steps / dt, "unit": UNIT,
                   "h2d_bytes_per_step": int(n_pairs * 24 + coords_h.numel() * 4),
                   "d2h_bytes_per_step": int(n_pairs * 32),
                   "api": "igmk_actdist_host_population (C ABI), pinned host buffers: the population (%.0f MB) is "
                          "staged again every step, as after an M-step, its upload pipelined with the pair "
                          "kernels" % pop_mb}
            if rank == 0:
                e2e["equal_to_two_calls"] = bool(torch.equal(h_out, ref_out))
            if world > 1:
                # N GPUs: the host side feeds N uploads of the same population through one
                # PCIe / memory system; replicate over NVLink instead
                coords_np_pinned = coords_h.numpy()
                dtr = timed("replicated")
                e2e_piped = e2e
                e2e = {"value": world * n_pairs * args.steps / dtr, "unit": UNIT,
                       "h2d_bytes_per_step": int(n_pairs * 24 + coords_h.numel() * 4 // world),
                       "d2h_bytes_per_step": int(n_pairs * 32),
                       "api": "igm_b200.dist.replicate_population (every rank uploads 1 / %d of the %.0f MB "
                              "population from pinned host memory, one all-gather over NVLink completes the "
                              "copies) + igmk_actdist_host (C ABI); the population is staged again every "
                              "step" % (world, pop_mb),
                       "per_rank_full_upload": {"value": e2e_piped["value"], "unit": UNIT,
                                                "api": e2e_piped["api"]}}
                if rank == 0:
                    e2e["equal_to_two_calls"] = bool(torch.equal(h_out, ref_out))
            e2e_two = {"value": world * n_pairs * args.steps / dt2, "unit": UNIT,
                       "h2d_bytes_per_step": e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                       "api": "igmk_upload_coords, then igmk_actdist_host (round 2's first definition)"}
        else:
            e2e = e2e_pairs
        sampler.mark_stop()      # the clock samples cover the timed regions

    # ---- the other configs of BASELINE.json, same run, outside the timed regions
    extra = {}
    if not args.no_extras and not args.max_pairs and args.nstruct == 1000 and args.resolution == 200_000:
        want = set(args.extras.split(","))
        if "4" in want:
            try:
                r4 = extra_config4(eng, nbead, args.nstruct, rank, world, dev, dist, torch, coords_np, radius)
                if rank == 0:
                    extra["config4_contact"] = r4
            except Exception as e:
                extra["config4_contact"] = {"error": repr(e)}
        if "3" in want:
            try:
                r3 = extra_config3(args, rank, world, dev, local_rank, dist, torch)
                if rank == 0:
                    extra["config3"] = r3
            except Exception as e:
                extra["config3"] = {"error": repr(e)}
        if "5" in want and world == 1:
            try:
                extra["config5"] = extra_config5(args, dev, local_rank, torch)
            except Exception as e:
                extra["config5"] = {"error": repr(e)}

    if rank == 0:
        time.sleep(0.1)
        sampler.stop()
        peak, peak_src = measured_peak_gbs()
        alg = algorithmic_bytes(ci, ii, jj, args.nstruct)
        k_ms = float(np.mean(kernel_ms))
        achieved = alg / (k_ms * 1e-3) / 1e9
        kname = "actdist_list_warp_kernel + key-array redo launch" if args.nstruct <= 1024 else \
            "actdist_list_block_kernel + key-array redo launch"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "%s: synthetic %d-structure population at %d kb male diploid "
                            "(%d beads), Hi-C A-step (%s) over the sigma=%g candidate list, "
                            "%d pairs per GPU" % (config_name(args.resolution, args.nstruct), args.nstruct,
                                                  args.resolution // 1000, nbead, args.mode, args.sigma, n_pairs),
                "pairs_per_gpu": n_pairs, "nstruct": args.nstruct, "nbead": nbead,
                "pair_structs_per_s": value * args.nstruct,
                "population": "seeded NumPy generator (same coordinates as --impl reference)" if host_pop
                              else "torch generator on the device",
                "parallelism": "pairs sharded over %d GPU(s), coordinates replicated, %s" % (
                    world, "finished records stored into every GPU's gather buffer from inside the kernel "
                    "(NVLink peer stores) + one barrier" if pg is not None else
                    "one NCCL all-gather of results"),
                "l2": "inputs larger than L2 (coordinates %.0f MB, pair list %.0f MB)" % (
                    nbead * 3 * eng.nstruct * 4 / 1e6, n_pairs * 24 / 1e6),
                "parity_sample_ok": parity,
                "gather_parity": gather_parity,
                "list_form_redo_pairs": redo_pairs,
                "sigma_sweep": sweep,
                "extra": extra,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic(args.nstruct, n_pairs, args.mode),
                         "kernel": kname, "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg,
                         "peak_source": peak_src},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        # what actually limits the kernel: the issue slots it uses and the
        # L1 / shared-memory data pipe, from the committed ncu capture rescaled to the live time
        winst = measured_traffic(args.nstruct, n_pairs, args.mode, "warp_instructions_per_launch")
        sm_mhz = line["clocks"].get("sm_mhz") or 1965.0
        if winst:
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            line["roofline"]["issue"] = {
                "warp_instructions_per_launch": winst,
                "achieved_winst_per_s": winst / (k_ms * 1e-3),
                "peak_winst_per_s": sms * 4 * sm_mhz * 1e6,
                "frac": winst / (k_ms * 1e-3) / (sms * 4 * sm_mhz * 1e6),
                "note": "issue slots used (4 schedulers per SM at the sampled SM clock)"}
        l1pct = measured_traffic(args.nstruct, n_pairs, args.mode, "l1tex_lsu_data_pipe_pct")
        ms_ncu = measured_traffic(args.nstruct, n_pairs, args.mode, "kernel_ms_under_ncu")
        if l1pct and ms_ncu:
            line["roofline"]["l1tex"] = {
                "lsu_data_pipe_frac": l1pct / 100.0 * ms_ncu / k_ms,
                "lsu_data_pipe_frac_under_ncu": l1pct / 100.0, "kernel_ms_under_ncu": ms_ncu,
                "note": "l1tex__data_pipe_lsu_wavefronts: coordinate loads, tile and list traffic"}
        if e2e is not None:
            line["e2e"] = e2e
            line["e2e_pairs_only"] = e2e_pairs
            if e2e_two is not None:
                line["e2e_two_calls"] = e2e_two
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
            # bounded CPU sample (about --cpu-seconds of work on all host cores): the reference's
            # own get_actdist on the leading slice of the same list, same population
            nb = min(n_pairs, 1_000_000)
            beads, remap, sub = sub_population(coords_get, ci, nbead, np.concatenate([ii[:nb], jj[:nb]]))

            class _CI2:
                def __getitem__(self, i):
                    return [int(remap[b]) for b in ci[i]]
            try:
                line["cpu_baseline"] = cpu_reference_rate(sub, radii[beads], chrom_hap, _CI2(), ii[:nb], jj[:nb],
                                                          pw[:nb], args.it_corr, args.mode, args.cpu_seconds)
            except Exception as e:                       # no copy of the reference on this box
                line["cpu_baseline"] = {"unavailable": repr(e)}
            # second CPU figure and a much larger parity gate: the plain-C oracle (OpenMP, all
            # host threads) over the first 1 M pairs, compared with the GPU results of the last
            # timed step pair by pair
            from oracle import c_oracle
            if c_oracle.available():
                n_c = int(nb)
                got_c = last[rank][:n_c].cpu().numpy().reshape(-1).view(_lib.PAIR_RESULT_DTYPE)
                equal, dt_c = c_oracle_check(ii[:n_c], jj[:n_c], pw[:n_c], sub, radii[beads], chrom_hap, ci, remap,
                                             args.it_corr, args.mode, got_c)
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
    """--impl reference: the reference's own CPU implementation of the path - its
    unmodified get_actdist (igm/steps/ActivationDistanceStep.py:336-485, imported from the
    reference tree or from the copy `make -C oracle` left in oracle/_ref) - on the host
    cores, SAME population (seeded NumPy generator) and SAME list as the GPU arm; each
    step is a bounded leading slice of that list.  Under torchrun only rank 0 works.
    The product library is not loaded by this arm."""
    if rank != 0:
        return
    from igm_b200 import synthetic
    chrom_hap, chrom_bead, copy_bead, ci, ii, jj, pw, _ = build_index_and_pairs(
        args.resolution, args.sigma, 0, args.max_pairs, native=False)
    nbead = len(chrom_bead)
    radius = float(synthetic.bead_radius(nbead))
    coords = host_population(chrom_bead, copy_bead, args.nstruct, radius)
    radii = np.full(nbead, radius, np.float32)
    steps = max(1, args.steps)
    budget = max(1.0, min(args.cpu_seconds, 120.0 / (steps + args.warmup)))
    kind = "reference"
    try:
        info = cpu_reference_rate(coords, radii, chrom_hap, ci, ii, jj, pw, args.it_corr, args.mode, budget,
                                  steps=steps, warmup=args.warmup)
    except Exception as e:
        sys.stderr.write("reference copy unavailable (%r); timing the oracle port instead\n" % (e,))
        info = port_rate(coords, radii, chrom_hap, ci, ii, jj, pw, args.it_corr, args.mode, budget, steps, args.warmup)
        kind = "port"
    value = info["value"]
    loaded = sorted(set(l.split()[-1] for l in open("/proc/self/maps") if ".so" in l and ROOT in l)) \
        if os.path.exists("/proc/self/maps") else []
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": info["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: synthetic %d-structure population at %d kb male diploid "
                               "(%d beads), Hi-C A-step (%s) over the sigma=%g candidate list; "
                               "bounded CPU sample per step (leading slice of the same list, same "
                               "population as the GPU arm)" % (config_name(args.resolution, args.nstruct),
                                                               args.nstruct, args.resolution // 1000,
                                                               nbead, args.mode, args.sigma),
                   "nstruct": args.nstruct, "nbead": nbead, "pairs_in_list": int(len(ii)),
                   "population": "seeded NumPy generator (same coordinates as the GPU arm)",
                   "repo_native_libraries_loaded": [os.path.relpath(p, ROOT) for p in loaded]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": kind,
                         "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _port_worker(sl):
    from oracle import actdist_oracle as orc
    ii, jj, pw, coords, radii, chrom_hap, ci, it_corr, mode = _G["p"]
    a, b = sl
    recs, _ = orc.run_pairs(ii[a:b], jj[a:b], pw[a:b], np.zeros(b - a), coords, radii, chrom_hap, ci, it_corr, 2.0, mode)
    return len(recs)


def port_rate(coords, radii, chrom_hap, ci, ii, jj, pw, it_corr, mode, budget_s, steps, warmup):
    """Fallback of the reference arm when neither /root/reference nor oracle/_ref/igm exists:
    the NumPy restatement (oracle/actdist_oracle.py), same pool / batch structure."""
    import multiprocessing as mp
    from oracle import actdist_oracle as orc
    cores = os.cpu_count() or 1
    _G["p"] = (ii, jj, pw, coords, radii, chrom_hap, ci, it_corr, orc.MODE_LB if mode == "LB" else orc.MODE_GP)
    t0 = time.perf_counter()
    _port_worker((0, min(200, len(ii))))
    rate1 = min(200, len(ii)) / (time.perf_counter() - t0)
    n = int(min(len(ii), max(1000, rate1 * cores * budget_s * 0.6)))
    batches = [(a, min(n, a + 1000)) for a in range(0, n, 1000)]
    rates, ms = [], []
    with mp.get_context("fork").Pool(cores) as pool:
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_port_worker, batches, chunksize=1)
            dt = time.perf_counter() - t0
            if k >= warmup:
                rates.append(n / dt)
                ms.append(dt * 1e3)
    return {"value": float(np.mean(rates)), "cores": cores, "ms_per_step": float(np.mean(ms)),
            "sample": "first %d candidate pairs, NumPy oracle port of get_actdist, fork pool of %d processes" % (n, cores)}


if __name__ == "__main__":
    main()
