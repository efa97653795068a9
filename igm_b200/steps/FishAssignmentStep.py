"""Drop-in replacement of IGM's ``FishAssignmentStep`` on B200
(igm/steps/FishAssignmentStep.py:83-389).

Same class name, ``name()`` string, config keys (``restraints/FISH/{input_fish, tol_list,
batch_size, fish_dir, keep_temporary_files}``, ``runtime/FISH/*``,
``optimization/structure_output``), batch layout ``(batch_id, 'pair' | 'probe', entries)``,
output ``fish_assignment.h5`` (``pairs, probes: int32``; ``pair_min, pair_max, radial_min,
radial_max: float32 (n, nstruct)``), swap-file naming and ``runtime/FISH/fish_assignment_file``.
``task`` hands the batch to the CUDA library (K5 ``rank_match_kernel``): per structure the
min / max distance over the copy combinations, its rank in the population, and the target
distance of that rank - ``target[np.argsort(np.argsort(.))]`` of :189-193, :214-219.

Deliberate differences:
* a pair's min / max run over ALL copy combinations, as the docstring of get_pair_dists
  says; the reference never advances its row counter (:33-40), so for more than one
  combination it takes the extrema over one real row and uninitialised memory;
* single-copy probes are accepted (the reference indexes ``ii[1]``, :52, and raises);
* ties are ranked by structure index (NumPy's default argsort order is unspecified there).
"""
from __future__ import annotations

import os
import shutil

import numpy as np

from .. import hdf5
from ..engine import ActdistEngine
from ..population import Population
from ._compat import Step, logger

_KEYS = ('pair_min', 'pair_max', 'radial_min', 'radial_max')


def _copies2(copy_index, loci) -> np.ndarray:
    """(n, 2) bead ids of the copies of each haploid locus, -1 where there is one copy."""
    ptr, beads = np.asarray(copy_index.ptr), np.asarray(copy_index.beads)
    loci = np.asarray(loci, dtype=np.int64)
    nc = ptr[loci + 1] - ptr[loci]
    if np.any(nc > 2):
        raise ValueError("at most two copies per locus are supported")
    out = np.full((len(loci), 2), -1, np.int32)
    out[:, 0] = beads[ptr[loci]]
    two = nc == 2
    out[two, 1] = beads[ptr[loci[two]] + 1]
    return out


class FishAssignmentStep(Step):

    def __init__(self, cfg):                                                          # :85-98
        if 'tol_list' not in cfg.get("runtime/FISH"):
            cfg["runtime"]["FISH"]["tol_list"] = cfg.get("restraints/FISH/tol_list")[:]
        if 'tol' not in cfg.get("runtime/FISH"):
            cfg["runtime"]["FISH"]["tol"] = cfg.get("runtime/FISH/tol_list").pop(0)
        super(FishAssignmentStep, self).__init__(cfg)

    def name(self):                                                                   # :101-109
        s = 'FishAssignmentStep (tol={:.2f}, iter={:s})'
        return s.format(self.cfg.get('runtime/FISH/tol', -1), str(self.cfg.get('runtime/opt_iter', 'N/A')))

    def setup(self):                                                                  # :112-155
        self.tmp_extensions = [".npz"]
        self.set_tmp_path()
        self.keep_temporary_files = self.cfg.get("restraints/FISH/keep_temporary_files", False)
        if not os.path.exists(self.tmp_dir):
            os.makedirs(self.tmp_dir)
        batch_size = self.cfg.get('restraints/FISH/batch_size')
        batches = []
        with hdf5.open_h5(self.cfg.get('restraints/FISH/input_fish')) as h5:
            entries = list(h5.keys())
            if 'pairs' in entries:
                logger.info('pairs are in FISH input!')
                pairs = np.asarray(h5['pairs'][()])
                for i in range(0, len(pairs), batch_size):
                    batches.append((len(batches), 'pair', pairs[i: i + batch_size]))
            if 'probes' in entries:
                logger.info('probes are in FISH input')
                probes = np.asarray(h5['probes'][()])
                for i in range(0, len(probes), batch_size):
                    batches.append((len(batches), 'probe', probes[i: i + batch_size]))
        self.argument_list = batches

    @staticmethod
    def task(batch, cfg, tmp_dir):                                                    # :158-236
        batch_id, entry_type, entries = batch
        pop = Population.from_hss(cfg.get("optimization/structure_output"))
        with hdf5.open_h5(cfg.get('restraints/FISH/input_fish')) as ftf:
            have = list(ftf.keys())
            tab = {k: np.asarray(ftf[k][()]) for k in have}
        out = {}
        with ActdistEngine(pop, int(cfg.get('restraints/FISH').get('gpu_device', 0))) as eng:
            if entry_type == 'pair' and len(entries):
                entries = np.asarray(entries).reshape(-1, 2)
                index = []
                for pair in entries:                                                  # :183 first matching row
                    hit = np.nonzero(np.all(tab['pairs'] == pair, axis=1))[0]
                    if len(hit) == 0:
                        raise ValueError("Cannot find pair: %s" % (pair,))
                    index.append(int(hit[0]))
                index = np.asarray(index, np.int64)
                a, b = _copies2(pop.copy_index, entries[:, 0]), _copies2(pop.copy_index, entries[:, 1])
                out['pair_index'] = index
                for key, red in (('pair_min', 'min'), ('pair_max', 'max')):
                    if key in have:
                        r = eng.rank_match(a, b, red, tab[key][index], want_rank=False, want_value=False)
                        out[key] = r['matched']
            if entry_type == 'probe' and len(entries):
                entries = np.asarray(entries).reshape(-1)
                index = []
                for probe in entries:                                                 # :205-210
                    hit = np.where(tab['probes'] == probe)[0]
                    if len(hit) != 1:
                        raise ValueError("Cannot find probe: %s" % (probe,))
                    index.append(int(hit[0]))
                index = np.asarray(index, np.int64)
                a = _copies2(pop.copy_index, entries)
                out['probe_index'] = index
                for key, red in (('radial_min', 'min'), ('radial_max', 'max')):
                    if key in have:
                        r = eng.rank_match(a, None, red, tab[key][index], want_rank=False, want_value=False)
                        out[key] = r['matched']
        np.savez(os.path.join(tmp_dir, 'tmp.%d.fish_targeting.npz' % batch_id), **out)

    def reduce(self):                                                                 # :239-332
        additional_data = []
        if "FISH" in self.cfg['runtime']:
            additional_data.append('tol_{:.4f}'.format(self.cfg['runtime']['FISH']['tol']))
        if 'opt_iter' in self.cfg['runtime']:
            additional_data.append('iter_{}'.format(self.cfg['runtime']['opt_iter']))
        fish_assignment_file = os.path.join(self.tmp_dir, "fish_assignment.h5")
        last_file = self.cfg['runtime']['FISH'].get("fish_assignment_file", None)
        with hdf5.open_h5(self.cfg.get('restraints/FISH/input_fish')) as h5:
            have = list(h5.keys())
            pairs = np.asarray(h5['pairs'][()]) if 'pairs' in have else None
            probes = np.asarray(h5['probes'][()]) if 'probes' in have else None
        nstruct = None
        rows = {k: {} for k in _KEYS}
        for batch_id, _, _ in self.argument_list:
            with np.load(os.path.join(self.tmp_dir, 'tmp.%d.fish_targeting.npz' % batch_id)) as t:
                for key in _KEYS:
                    if key in t.files:
                        idx = t['pair_index' if key.startswith('pair') else 'probe_index']
                        for k, v in zip(idx.tolist(), t[key]):
                            rows[key][k] = v
                            nstruct = len(v)
        data = {}
        if pairs is not None:
            data['pairs'] = pairs.astype(np.int32)
        if probes is not None:
            data['probes'] = probes.astype(np.int32)
        for key in _KEYS:
            if key in have:
                n = len(pairs) if key.startswith('pair') else len(probes)
                missing = [k for k in range(n) if k not in rows[key]]
                if missing:
                    raise ValueError("%s: no assignment for entries %s" % (key, missing[:5]))
                data[key] = (np.stack([rows[key][k] for k in range(n)]).astype(np.float32)
                             if n else np.zeros((0, nstruct or 0), np.float32))
        tmp_file = fish_assignment_file + '.tmp'
        hdf5.write_h5(tmp_file, data)
        swapfile = os.path.realpath('.'.join([fish_assignment_file, ] + additional_data))
        if last_file is not None:
            shutil.move(last_file, swapfile)
        shutil.move(tmp_file, fish_assignment_file)
        self.cfg['runtime']['FISH']["fish_assignment_file"] = fish_assignment_file

    def skip(self):                                                                   # :365-373
        self.set_tmp_path()
        self.cfg['runtime']['FISH']["fish_assignment_file"] = os.path.join(self.tmp_dir, "fish_assignment.h5")

    def set_tmp_path(self):                                                           # :376-387
        fish_tmp_dir = self.cfg['restraints']['FISH'].get('fish_dir', 'fish_actdist')
        if os.path.isabs(fish_tmp_dir):
            self.tmp_dir = fish_tmp_dir
        else:
            self.tmp_dir = os.path.abspath(os.path.join(self.cfg['parameters']['tmp_dir'], fish_tmp_dir))
