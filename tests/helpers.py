"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

from igm_b200.population import CopyIndex, Population

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
REF_DEMO = os.path.join(ROOT, "oracle", "_ref", "demo")
DEMO_HSS = os.path.join(REF_DEMO, "igm-model.hss.T")
DEMO_HCS = os.path.join(REF_DEMO, "WTC11_HiC_2Mb.hcs")


def have_demo():
    return os.path.exists(DEMO_HSS) and os.path.exists(DEMO_HCS)


def load_case(npz, prefix):
    pop = Population(npz[prefix + "_coords"], npz[prefix + "_radii"], npz[prefix + "_chrom"],
                     CopyIndex(npz[prefix + "_copy_ptr"], npz[prefix + "_copy_beads"]))
    return (pop, npz[prefix + "_i"], npz[prefix + "_j"], npz[prefix + "_pwish"],
            npz[prefix + "_plast"])


def golden_out(npz, prefix, mode, it_corr):
    key = "%s_%s_it%d" % (prefix, mode, it_corr)
    return {k: npz[key + "_" + k] for k in ("row", "col", "ad", "p", "nrec")} | {
        "sha": str(npz[key + "_sha"])}


def demo_full_summary():
    with open(os.path.join(GOLDEN, "demo_full.json")) as f:
        return json.load(f)


def all_small_cases():
    """Yields (name, npz, prefix) for every committed small golden case."""
    d = np.load(os.path.join(GOLDEN, "demo_subset.npz"))
    yield "demo", d, "demo"
    s = np.load(os.path.join(GOLDEN, "synth_small.npz"))
    for name in s["names"]:
        yield str(name), s, str(name)


def sprite_golden_cases():
    """Yields (name, coords, chrom_per_bead, copy_index dict, clusters, rg2s, best, selected list, seed0)
    for tests/golden/sprite_small.npz (populations of damid_small.npz)."""
    g = np.load(os.path.join(GOLDEN, "sprite_small.npz"))
    d = np.load(os.path.join(GOLDEN, "damid_small.npz"))
    for name in [str(x) for x in g["names"]]:
        crd = np.ascontiguousarray(d[name + "_coords"], np.float32)
        ptr, beads = d[name + "_copy_ptr"], d[name + "_copy_beads"]
        copy_index = {i: [int(b) for b in beads[ptr[i]:ptr[i + 1]]] for i in range(len(ptr) - 1)}
        cp, cd = g[name + "_cluster_ptr"], g[name + "_cluster_data"]
        clusters = [cd[cp[k]:cp[k + 1]] for k in range(len(cp) - 1)]
        nstruct = crd.shape[1]
        sel, off = [], 0
        for cl in clusters:
            sel.append(g[name + "_selected"][off:off + nstruct * len(cl)].reshape(nstruct, len(cl)))
            off += nstruct * len(cl)
        yield (name, crd, np.asarray(d[name + "_chrom"]), copy_index, clusters, g[name + "_rg2s"],
               g[name + "_best"], sel, int(g["seed0"]), (ptr, beads))


def sprite_reference_task(crd, chrom, copy_index, clusters, keep_best, max_chrom):
    """Restatement of SpriteAssignmentStep.task for one batch (reference :103-152) on top of
    the pinned oracle port of compute_gyration_radius.  Returns (selected, indexes, values)."""
    from oracle import sprite_oracle as so
    indexes, values, selected = [], [], []
    for cluster in clusters:
        n_chrom = len(np.unique(np.asarray(chrom)[cluster]))
        if n_chrom > max_chrom:
            selected.append(np.zeros((keep_best, len(cluster)), dtype='i4') - 1)
            indexes.append(np.array([-1] * keep_best))
            values.append(np.array([-1] * keep_best))
            continue
        rg2s, _, cur = so.compute_gyration_radius_port(crd, cluster, chrom, copy_index)
        ind = np.argpartition(rg2s, keep_best)[:keep_best]
        ind = ind[np.argsort(rg2s[ind])]
        selected.append(cur[ind])
        indexes.append(ind)
        values.append(rg2s[ind])
    return selected, np.array(indexes, dtype=np.int32), np.asarray(values)


def sprite_reference_reduce(batches, indptr, n_struct, batch_size, kT):
    """Restatement of SpriteAssignmentStep.reduce (reference :166-259): ``batches`` is the
    list of (selected, indexes, values) per batch id.  Consumes np.random like the reference."""
    n_clusters = len(indptr) - 1
    random_order = np.random.permutation(range(len(batches)))
    occupancy = np.zeros(n_struct, dtype=np.int32)
    assignment = np.zeros(n_clusters, dtype=np.int32)
    aveN = float(n_clusters) / n_struct
    stdN = np.sqrt(aveN)
    selected = np.zeros(int(indptr[-1]), np.int32)
    for batch_id in random_order:
        sel, idx, val = batches[batch_id]
        assigned = []
        for i, (best_rg2s, curr_idx) in enumerate(zip(val, idx)):
            ci = i + batch_id * batch_size
            if best_rg2s[0] < 0:
                pos, si = 0, -1
            else:
                best_rgs = np.sqrt(best_rg2s)
                pen = np.clip(occupancy[curr_idx] - aveN, 0., None) / stdN
                E = (best_rgs - best_rgs[0]) / kT + pen
                P = np.cumsum(np.exp(-(E - E[0])))
                e = np.random.rand() * P[-1]
                pos = np.searchsorted(P, e, side='left')
                si = curr_idx[pos]
                occupancy[si] += 1
            assignment[ci] = si
            assigned.append(sel[i][pos])
        start = indptr[batch_id * batch_size]
        stop = indptr[batch_id * batch_size + len(assigned)]
        selected[start:stop] = np.concatenate(assigned)
    return assignment, selected
