"""Drop-in replacement of IGM's ``PolymerAssignmentStep`` on B200
(igm/steps/PolymerAssignmentStep.py:36-204).

Same class name, ``name()`` string, config keys (``restraints/polymer/{polymer_file,
assignment_file}``, ``runtime/polymer/{tmp_dir, assignment_file}``,
``optimization/structure_output``), batches of 1000 consecutive beads ``(batch_id, range)``,
output datasets ``loci: int32`` and ``nn_dist: float32 (n_bonds, nstruct)`` and swap-file
naming.  Per bead i the reference draws nstruct distances from the experimental histogram
(``np.sort(np.random.choice(edges, nstruct, p=probability))``, :115 - the same NumPy calls, in
the same order, are made here so a seeded run draws the same samples) and hands the
structure whose (i, i+1) distance has rank k the k-th sample; the distances, ranks and
the gather run on the GPU (K5 ``rank_match_kernel``).  Ties are ranked by structure index.

The reference's last batch contains bead nbead-1 when ``nbead % 1000 == 0`` (:76-80) and
then indexes past the coordinates; here the batch is cut at the last bond.
"""
from __future__ import annotations

import os
import shutil

import numpy as np

from .. import hdf5
from ..engine import ActdistEngine
from ..population import Population
from ._compat import Step, logger


class PolymerAssignmentStep(Step):

    def __init__(self, cfg):                                                          # :38-41
        super(PolymerAssignmentStep, self).__init__(cfg)

    def name(self):                                                                   # :44-51
        return 'PolymerAssignmentStep (iter={:s})'.format(str(self.cfg.get('runtime/opt_iter', 'N/A')))

    def setup(self):                                                                  # :54-82
        self.tmp_extensions = [".npz"]
        self.set_tmp_path()
        if not os.path.exists(self.tmp_dir):
            os.makedirs(self.tmp_dir)
        with hdf5.open_h5(self.cfg.get("optimization/structure_output")) as h5:
            nbead = int(h5['coordinates'].shape[0])
        batch_size = 1000
        batches = []
        for i in range(0, nbead - 1, batch_size):
            batches.append((len(batches), range(i, min(i + batch_size, nbead - 1))))
        self.argument_list = batches

    @staticmethod
    def task(batch, cfg, tmp_dir):                                                    # :84-126
        batch_id, entries = batch
        pop = Population.from_hss(cfg.get("optimization/structure_output"))
        with hdf5.open_h5(cfg.get('restraints/polymer/polymer_file')) as ftf:
            edges = np.asarray(ftf['bin_edges'][()])
            dist_prob = np.asarray(ftf['probability'][()])
        loci = np.asarray(list(entries), np.int32)
        nstruct = pop.nstruct
        target = np.empty((len(loci), nstruct), np.float32)
        for k in range(len(loci)):                                                    # :115, one draw per bead, in bead order
            target[k] = np.sort(np.random.choice(edges, nstruct, p=dist_prob))
        a = np.stack([loci, np.full(len(loci), -1, np.int32)], axis=1)
        b = np.stack([loci + 1, np.full(len(loci), -1, np.int32)], axis=1)
        with ActdistEngine(pop, int(cfg.get('restraints/polymer').get('gpu_device', 0))) as eng:
            r = eng.rank_match(a, b, "min", target, want_rank=False, want_value=False)
        np.savez(os.path.join(tmp_dir, 'tmp.%d.polymer.npz' % batch_id), nn_dist=r['matched'])

    def reduce(self):                                                                 # :129-181
        additional_data = []
        if 'opt_iter' in self.cfg['runtime']:
            additional_data.append('iter_{}'.format(self.cfg['runtime']['opt_iter']))
        polymer_assignment_file = os.path.join(self.tmp_dir, self.cfg['restraints']['polymer']['assignment_file'])
        last_file = self.cfg['runtime']['polymer'].get("assignment_file", None)
        beads, nn_dist = [], []
        for batch_id, entries in self.argument_list:
            with np.load(os.path.join(self.tmp_dir, 'tmp.%d.polymer.npz' % batch_id)) as t:
                nn_dist.append(t['nn_dist'])
            beads.append(np.asarray(list(entries), np.int32))
        tmp_file = polymer_assignment_file + '.tmp'
        logger.info(polymer_assignment_file)
        hdf5.write_h5(tmp_file, {'loci': np.concatenate(beads).astype(np.int32),
                                 'nn_dist': np.concatenate(nn_dist).astype(np.float32)})
        swapfile = os.path.realpath('.'.join([polymer_assignment_file, ] + additional_data))
        if last_file is not None:
            shutil.move(last_file, swapfile)
        shutil.move(tmp_file, polymer_assignment_file)
        self.cfg['runtime']['polymer']["assignment_file"] = polymer_assignment_file

    def skip(self):                                                                   # :184-191
        self.set_tmp_path()
        self.cfg['runtime']['polymer']["assignment_file"] = os.path.join(self.tmp_dir, "polymer_assignment.h5")

    def set_tmp_path(self):                                                           # :194-204
        poly_tmp_dir = self.cfg['runtime']['polymer'].get('tmp_dir', 'poly_actdist')
        if os.path.isabs(poly_tmp_dir):
            self.tmp_dir = poly_tmp_dir
        else:
            self.tmp_dir = os.path.abspath(os.path.join(self.cfg['parameters']['tmp_dir'], poly_tmp_dir))
