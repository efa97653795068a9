"""Drop-in replacement of IGM's lamina ``DamidActivationDistanceStep`` on B200
(igm/steps/DamidActivationDistanceStep.py:41-359), spherical envelope.

Same class name, ``name()`` string, config keys (``restraints/DamID/{input_profile,
sigma_list, contact_range, tmp_dir, keep_temporary_files, batch_size}``,
``model/restraints/envelope/{nucleus_shape, nucleus_radius}``,
``optimization/{structure_output, iter_corr_knob}``, ``runtime/DamID/*``), temp-file
names (``%d.damid.in.npy`` float32 batches), output ``damid_actdist.hdf5``
(``loc: int32``, ``dist, prob: float32``) and swap-file naming.  ``task`` hands the
batch to the CUDA library (the Hi-C select kernel with an all-zero partner row);
the 5-decimal text round trip of the reference (:35, :270, :289) is reproduced on
the device.  ``exp_map`` envelopes (volume files) and ellipsoids are not
supported - the reference itself never evaluates ellipsoids (:245).
"""
from __future__ import annotations

import os
import shutil

import numpy as np

from .. import hdf5
from ..engine import ActdistEngine
from ..population import Population
from ._compat import Step, logger, make_absolute_path

damid_actdist_shape = [('loc', 'int32'), ('dist', 'float32'), ('prob', 'float32')]   # :30-34
damid_actdist_fmt_str = "%6d %.5f %.5f"                                              # :35


class DamidActivationDistanceStep(Step):

    # The nuclear-body variant (NuclDamidActivationDistanceStep.py) is the same arithmetic
    # under other config keys; the keys are class attributes so it can subclass this step.
    _KEY = "DamID"                      # restraints/<KEY>, runtime/<KEY>
    _BODY = "envelope"                  # model/restraints/<BODY>/{nucleus_shape, nucleus_radius}
    _STEP_NAME = 'DamidActivationDistanceStep'
    _TMP_DEFAULT = 'damid_actdist'      # setup (:100 / Nucl :172)
    _TMP_DEFAULT_SKIP = 'damid_actdist'  # skip (:356 in both files)

    def __init__(self, cfg):                                                          # :43-67
        K = self._KEY
        rt = cfg["runtime"].setdefault(K, {})
        if 'sigma_list' not in rt:
            rt["sigma_list"] = cfg.get("restraints/%s/sigma_list" % K)[:]
        if "sigma" not in rt:
            rt["sigma"] = rt["sigma_list"].pop(0)
        if "iter_corr_knob" not in rt:
            rt["iter_corr_knob"] = cfg.get("optimization/iter_corr_knob")
        if rt["iter_corr_knob"] == 1:
            logger.info('Using iterative correction (standard choice)')
        else:
            logger.info('Not using iterative correction (number of contacts may blow up!)')
        Step.__init__(self, cfg)

    def name(self):                                                                   # :69-77
        s = self._STEP_NAME + ' (sigma={:.2f}%, iter={:s})'
        return s.format(self.cfg.get('runtime/%s/sigma' % self._KEY, -1) * 100.0,
                        str(self.cfg.get('runtime/opt_iter', 'N/A')))

    def _tmp_dir(self, default=None):
        return make_absolute_path(self.cfg.get('restraints/%s/tmp_dir' % self._KEY, default or self._TMP_DEFAULT),
                                  self.cfg.get('parameters/tmp_dir'))

    def setup(self):                                                                  # :80-150
        K = self._KEY
        sigma = self.cfg.get("runtime/%s/sigma" % K)
        input_profile = self.cfg.get("restraints/%s/input_profile" % K)
        last_file = self.cfg.get('runtime/%s' % K).get("damid_actdist_file", None)
        batch_size = self.cfg.get('restraints/%s/batch_size' % K, 100)
        self.tmp_extensions = [".npy", ".tmp"]
        self.tmp_dir = self._tmp_dir()
        if not os.path.exists(self.tmp_dir):
            os.makedirs(self.tmp_dir)
        self.keep_temporary_files = self.cfg.get("restraints/%s/keep_temporary_files" % K, False)
        profile = np.loadtxt(input_profile, dtype='float32')
        mask = profile >= sigma
        ii = np.where(mask)[0]
        p_exp = profile[mask]
        plast = np.zeros(len(ii), dtype=np.float32)
        if last_file is not None:
            with hdf5.open_h5(last_file) as h5f:
                loc = np.asarray(h5f["loc"][()]).astype(np.int64)
                prob = np.asarray(h5f["prob"][()]).astype(np.float32)
            # dict(zip(loc, prob)): the LAST record of a locus wins (:129)
            if len(loc):
                table = np.zeros(int(max(loc.max(), ii.max() if len(ii) else 0)) + 1, np.float32)
                table[loc] = prob
                plast = table[ii]
        n_args_batches = len(ii) // batch_size + (1 if len(ii) % batch_size else 0)
        for b in range(n_args_batches):
            start, end = b * batch_size, min((b + 1) * batch_size, len(ii))
            params = np.stack([ii[start:end].astype(np.float32), p_exp[start:end], plast[start:end]], axis=1)
            np.save(os.path.join(self.tmp_dir, '%d.damid.in.npy' % b), params.astype(np.float32))
        self.argument_list = range(n_args_batches)

    @classmethod
    def task(cls, batch_id, cfg, tmp_dir):                                            # :152-270
        K, B = cls._KEY, cls._BODY
        shape = cfg.get('model/restraints/%s/nucleus_shape' % B)
        if shape != 'sphere':
            raise NotImplementedError('%s restraint for shape %s has not been implemented on the GPU path.' % (K, shape))
        radius = cfg.get('model/restraints/%s/nucleus_radius' % B)
        it_corr = 1 if cfg.get('runtime/%s/iter_corr_knob' % K) == 1 else 0
        params = np.load(os.path.join(tmp_dir, '%d.damid.in.npy' % batch_id)).reshape(-1, 3)
        pop = Population.from_hss(cfg.get("optimization/structure_output"))
        with ActdistEngine(pop, int(cfg.get('restraints/%s' % K).get('gpu_device', 0))) as eng:
            loci = params[:, 0].astype(np.int32)
            res = eng.damid_actdist(loci, params[:, 1], params[:, 2], float(radius),
                                    cfg.get('restraints/%s/contact_range' % K, 0.05), it_corr)
            loc, dist, prob = eng.expand_damid_records(loci, res, pop.copy_index.ptr, pop.copy_index.beads)
        np.savez(os.path.join(tmp_dir, '%d.damid.out.npz' % batch_id), loc=loc, dist=dist, prob=prob)
        if cfg.get('restraints/%s' % K).get('write_text_tmp', False):
            rep = np.repeat(np.arange(len(loci)), res["nrec"])
            ad = np.where(res["o"] >= 0,
                          np.sqrt(res["d2_sel_bits"].view(np.float32).astype(np.float64) / _denoms(pop, loci, radius, cfg, K)),
                          2.0)[rep]
            with open(os.path.join(tmp_dir, '%d.out.tmp' % batch_id), 'w') as f:
                f.write('\n'.join([damid_actdist_fmt_str % x for x in zip(loc.tolist(), ad.tolist(), res["p"][rep].tolist())]))

    def reduce(self):                                                                 # :272-337
        damid_actdist_file = os.path.join(self.tmp_dir, "damid_actdist.hdf5")
        K = self._KEY
        last_file = self.cfg['runtime'][K].get("damid_actdist_file", None)
        loc, dist, prob = [np.zeros(0, np.int32)], [np.zeros(0, np.float32)], [np.zeros(0, np.float32)]
        for i in self.argument_list:
            with np.load(os.path.join(self.tmp_dir, '%d.damid.out.npz' % i)) as z:
                loc.append(z['loc']); dist.append(z['dist']); prob.append(z['prob'])
        additional_data = []
        if K in self.cfg['runtime']:
            additional_data.append((K + '_{:.4f}').format(self.cfg['runtime'][K]['sigma']))
        if 'opt_iter' in self.cfg['runtime']:
            additional_data.append('iter_{}'.format(self.cfg['runtime']['opt_iter'] - 1))
        tmp_file = damid_actdist_file + '.tmp'
        hdf5.write_h5(tmp_file, {"loc": np.concatenate(loc).astype(np.int32),
                                 "dist": np.concatenate(dist).astype(np.float32),
                                 "prob": np.concatenate(prob).astype(np.float32)})
        swapfile = os.path.realpath('.'.join([damid_actdist_file, ] + additional_data))
        if last_file is not None:
            shutil.move(last_file, swapfile)
        shutil.move(tmp_file, damid_actdist_file)
        self.cfg['runtime'][K]["damid_actdist_file"] = damid_actdist_file

    def cleanup(self):
        if not self.keep_temporary_files:
            for f in os.listdir(self.tmp_dir):
                if f.endswith('.damid.out.npz'):
                    os.remove(os.path.join(self.tmp_dir, f))
        Step.cleanup(self)

    def skip(self):                                                                   # :339-359
        self.tmp_dir = self._tmp_dir(self._TMP_DEFAULT_SKIP)
        self.damid_actdist_file = os.path.join(self.tmp_dir, "damid_actdist.hdf5")
        self.cfg['runtime'][self._KEY]["damid_actdist_file"] = self.damid_actdist_file


class NuclDamidActivationDistanceStep(DamidActivationDistanceStep):
    """Nuclear-body DamID A-step (igm/steps/NuclDamidActivationDistanceStep.py:120-361):
    the lamina step's arithmetic under ``restraints/nuclDamID``, ``runtime/nuclDamID`` and
    ``model/restraints/nucleolus`` (sphere).  As in the reference, ``setup`` defaults the
    directory to ``nucldamid_actdist`` (:172) while ``skip`` defaults it to
    ``damid_actdist`` (:356)."""
    _KEY = "nuclDamID"
    _BODY = "nucleolus"
    _STEP_NAME = 'NuclDamidActivationDistanceStep'
    _TMP_DEFAULT = 'nucldamid_actdist'
    _TMP_DEFAULT_SKIP = 'damid_actdist'


def _denoms(pop, loci, radius, cfg, key="DamID"):
    """(R - r)^2 per locus in float64, as the reference evaluates it (:436,:441)."""
    R = np.array(radius) * (1 - cfg.get('restraints/%s/contact_range' % key, 0.05))
    r = pop.radii[np.asarray(pop.copy_index.beads)[np.asarray(pop.copy_index.ptr)[loci]]]
    return (R - r.astype(np.float64)) ** 2
