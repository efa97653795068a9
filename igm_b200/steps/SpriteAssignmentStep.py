"""Drop-in replacement of IGM's ``SpriteAssignmentStep`` on B200
(igm/steps/SpriteAssignmentStep.py:19-259; arithmetic in
igm/cython_compiled/sprite.pyx:104-283 and cpp_sprite_assignment.cpp:79-143).

Same class name, ``name()`` string, config keys (``restraints/sprite/{clusters,
volume_fraction_list, batch_size, keep_best, max_chrom_in_cluster, radius_kt, tmp_dir,
assignment_file, keep_temporary_files}``, ``runtime/sprite/*``,
``optimization/structure_output``), batch layout (``batch_size`` clusters per task), temp files
(``tmp.%d.selected.npz``, ``tmp.%d.idx.npy``, ``tmp.%d.values.npy``) and output
``assignment.h5`` (``assignment: int32 (n_clusters,)``, ``selected: int32 (indptr[-1],)``,
``indptr``).

``task`` evaluates a whole batch of clusters on the GPU: K4 (``sprite_rg2_kernel``, the
reference's get_rg2s_cpp) picks, per structure, the chromosome copies that minimise the Rg of
one representative segment per chromosome; K4b (``sprite_cluster_rg2_kernel``) then computes
the Rg^2 of the full cluster under that choice, straight from the population resident in HBM.
The random draws stay on the host and consume NumPy's global stream exactly as the reference
does (``np.random.choice`` per chromosome in ``task``, ``np.random.permutation`` and one
``np.random.rand`` per cluster in ``reduce``), so a seeded run reproduces the reference's
assignment.
"""
from __future__ import annotations

import os

import numpy as np

from .. import hdf5
from ._compat import Step, make_absolute_path
from .ActivationDistanceStep import _get_engine


def batch_gyration_radii(eng, chrom, copy_index, clusters, max_chrom, choice=None):
    """compute_gyration_radius (sprite.pyx:104-283) for a list of clusters on the GPU.

    Returns one entry per cluster: ``None`` when the cluster spans more than ``max_chrom``
    chromosomes (SpriteAssignmentStep.py:118-125), else ``(rg2s, select)`` where
    ``select(ind)`` gives the selected bead ids ``(len(ind), n_segments)`` of the
    structures ``ind``."""
    choice = choice or np.random.choice
    nstruct = eng.nstruct
    plans = []          # per cluster: None | dict
    k4_regions = []     # stage one: the copies of one representative segment per chromosome
    full = []           # stage two / single-chromosome variants: (segments, groups, sel)
    for cluster in clusters:
        cl = np.sort(np.asarray(cluster))                                        # pyx :193
        cchroms = np.asarray(chrom)[cl]
        uniq = np.unique(cchroms)
        if len(uniq) > max_chrom:
            plans.append(None)
        elif len(uniq) == 1:                                                     # pyx :198-215
            nc = len(copy_index[int(cl[0])])
            bead_group = np.array([[copy_index[int(i)][k] for i in cl] for k in range(nc)], np.int32)
            plans.append({"kind": "single", "first": len(full), "nc": nc, "bead_group": bead_group})
            for k in range(nc):                      # one constant selection per copy
                full.append(([[int(b)] for b in bead_group[k]], [0] * len(cl),
                             np.zeros((nstruct, 1), np.int32)))
        else:
            segs = [cl[cchroms == c] for c in uniq]                              # pyx :218-220
            reps = [int(choice(x)) for x in segs if len(x)]                      # pyx :222-223 (global RNG)
            plans.append({"kind": "multi", "k4": len(k4_regions), "segs": segs})
            k4_regions.append([copy_index[r] for r in reps])
    stage1 = eng.sprite_rg2(k4_regions) if k4_regions else []
    for p in plans:
        if p is not None and p["kind"] == "multi":
            p["sel"] = stage1[p["k4"]][2]                                        # (nstruct, n_representatives)
            p["all_segments"] = np.concatenate(p["segs"])                        # pyx :253
            p["groups"] = np.concatenate([np.full(len(x), g, np.int64) for g, x in enumerate(p["segs"])])
            p["full"] = len(full)
            full.append(([copy_index[int(i)] for i in p["all_segments"]], p["groups"].tolist(), p["sel"]))
    rg_full = eng.sprite_cluster_rg2(full)
    out = []
    for p in plans:
        if p is None:
            out.append(None)
        elif p["kind"] == "single":
            rgs = rg_full[p["first"]:p["first"] + p["nc"]]
            cidx = np.argsort(rgs, axis=0)[0, :]                                 # pyx :212
            rg = rgs[cidx, np.arange(nstruct)]
            out.append((rg, (lambda ind, bg=p["bead_group"], ci=cidx: bg[ci[np.asarray(ind)]])))
        else:
            loc = [np.asarray(copy_index[int(i)]) for i in p["all_segments"]]

            def select(ind, loc=loc, groups=p["groups"], sel=p["sel"]):          # pyx :270-272
                s = sel[np.asarray(ind)]                                         # (len(ind), n_groups)
                return np.stack([loc[t][s[:, groups[t]]] for t in range(len(loc))], axis=1).astype(np.int32)
            out.append((rg_full[p["full"]], select))
    return out


class SpriteAssignmentStep(Step):

    def __init__(self, cfg):                                                      # :21-33
        if 'volume_fraction_list' not in cfg.get("runtime/sprite"):
            cfg["runtime"]["sprite"]["volume_fraction_list"] = cfg.get("restraints/sprite/volume_fraction_list")[:]
        if 'volume_fraction' not in cfg.get("runtime/sprite"):
            cfg["runtime"]["sprite"]["volume_fraction"] = cfg.get("runtime/sprite/volume_fraction_list").pop(0)
        super(SpriteAssignmentStep, self).__init__(cfg)

    def name(self):                                                               # :36-44
        s = 'SpriteAssignmentStep (volume_fraction={:.1f}%, iter={:s})'
        return s.format(self.cfg.get('runtime/sprite/volume_fraction', -1),
                        str(self.cfg.get('runtime/opt_iter', 'N/A')))

    def setup(self):                                                              # :46-79
        self.tmp_extensions = [".npy", ".npz"]
        self.tmp_dir = make_absolute_path(self.cfg.get('restraints/sprite/tmp_dir', 'sprite'),
                                          self.cfg.get('parameters/tmp_dir'))
        self.keep_temporary_files = self.cfg.get('restraints/sprite/keep_temporary_files', False)
        if not os.path.isdir(self.tmp_dir):
            os.makedirs(self.tmp_dir)
        with hdf5.open_h5(self.cfg.get('restraints/sprite/clusters')) as h5:
            n_clusters = len(h5['indptr']) - 1
        with hdf5.open_h5(self.cfg.get("optimization/structure_output")) as hss:
            self.n_struct = int(hss["coordinates"].shape[1])
        batch_size = self.cfg.get('restraints/sprite/batch_size', 10)
        n_batches = n_clusters // batch_size + (1 if n_clusters % batch_size else 0)
        self.n_batches = n_batches
        self.n_clusters = n_clusters
        self.argument_list = range(n_batches)

    @staticmethod
    def task(batch_id, cfg, tmp_dir):                                             # :81-164
        batch_size = cfg.get('restraints/sprite/batch_size', 10)
        keep_best = cfg.get('restraints/sprite/keep_best', 50)
        max_chrom = cfg.get('restraints/sprite/max_chrom_in_cluster', 6)
        with hdf5.open_h5(cfg.get('restraints/sprite/clusters')) as h5:
            indptr = np.asarray(h5['indptr'][()])
            ii = indptr[batch_id * batch_size:(batch_id + 1) * batch_size + 1].astype(np.int64)
            data = np.asarray(h5['data'][()])[ii[0]:ii[-1]]
        ii = ii - ii[0]
        clusters = [data[ii[i - 1]:ii[i]] for i in range(1, len(ii))]
        eng = _get_engine(cfg.get("optimization/structure_output"),
                          int(cfg.get('restraints/sprite').get('gpu_device', 0)))
        results = batch_gyration_radii(eng, eng.chrom, eng.copy_index, clusters, max_chrom)
        indexes, values, selected_beads = [], [], []
        for cluster, res in zip(clusters, results):
            if res is None:                                                       # :118-125
                selected_beads.append(np.zeros((keep_best, len(cluster)), dtype='i4') - 1)
                indexes.append(np.array([-1] * keep_best))
                values.append(np.array([-1] * keep_best))
                continue
            rg2s, select = res
            ind = np.argpartition(rg2s, keep_best)[:keep_best]                    # :143-144
            ind = ind[np.argsort(rg2s[ind])]
            selected_beads.append(select(ind))
            indexes.append(ind)
            values.append(rg2s[ind])
        np.savez(os.path.join(tmp_dir, 'tmp.%d.selected.npz' % batch_id), *selected_beads)
        np.save(os.path.join(tmp_dir, 'tmp.%d.idx.npy' % batch_id), np.array(indexes, dtype=np.int32))
        np.save(os.path.join(tmp_dir, 'tmp.%d.values.npy' % batch_id), values)

    def reduce(self):                                                             # :166-259
        random_order = np.random.permutation(self.argument_list)
        batch_size = self.cfg.get('restraints/sprite/batch_size', 10)
        kT = self.cfg.get('restraints/sprite/radius_kt', 100.0)
        occupancy = np.zeros(self.n_struct, dtype=np.int32)
        assignment = np.zeros(self.n_clusters, dtype=np.int32)
        aveN = float(self.n_clusters) / self.n_struct
        stdN = np.sqrt(aveN)
        with hdf5.open_h5(self.cfg.get('restraints/sprite/clusters')) as h5:
            indptr = np.asarray(h5['indptr'][()])
        assignment_filename = make_absolute_path(
            self.cfg.get('restraints/sprite/assignment_file', 'assignment.h5'), self.tmp_dir)
        selected = np.zeros(int(indptr[-1]), dtype=np.int32)
        for batch_id in random_order:
            structure_indexes = np.load(os.path.join(self.tmp_dir, 'tmp.%d.idx.npy' % batch_id))
            rg2_values = np.load(os.path.join(self.tmp_dir, 'tmp.%d.values.npy' % batch_id))
            selected_beads_zip = np.load(os.path.join(self.tmp_dir, 'tmp.%d.selected.npz' % batch_id))
            assigned_beads = []
            for i, (best_rg2s, curr_idx) in enumerate(zip(rg2_values, structure_indexes)):
                ci = i + batch_id * batch_size
                if best_rg2s[0] < 0:                                              # cluster was skipped in task
                    pos, si = 0, -1
                else:
                    # Gibbs draw over the keep_best most compact structures, penalising
                    # structures that already express more than their share (:218-238)
                    best_rgs = np.sqrt(best_rg2s)
                    penal = np.clip(occupancy[curr_idx] - aveN, 0., None) / stdN
                    E = (best_rgs - best_rgs[0]) / kT + penal
                    P = np.cumsum(np.exp(-(E - E[0])))
                    e = np.random.rand() * P[-1]
                    pos = np.searchsorted(P, e, side='left')
                    si = curr_idx[pos]
                    occupancy[si] += 1
                assignment[ci] = si
                assigned_beads.append(selected_beads_zip['arr_%i' % i][pos])
            start = indptr[batch_id * batch_size]
            stop = indptr[batch_id * batch_size + len(assigned_beads)]
            selected[start:stop] = np.concatenate(assigned_beads)
        hdf5.write_h5(assignment_filename, {"assignment": assignment, "selected": selected,
                                            "indptr": indptr})
