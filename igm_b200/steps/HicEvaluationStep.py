"""Drop-in for the contact-map part of IGM's ``HicEvaluationStep``
(igm/steps/HicEvaluationStep.py:19-179) on B200.

Same class name, ``name()`` format, ``setup`` (empty ``argument_list``, output
directory ``<workdir>/evaluation/Hi-C/sigma_XX.iter_YY``) and ``reduce``
products that are plain data: ``out_matrix.hcs`` (copies summed, clipped to
[0, 1], :109-113) and ``stats.txt`` (:156-176).  The matplotlib plots of the
reference (:114-155) and the bead-level ``full_matrix.hcs`` are out of scope
(SURVEY.md section 2 row 3).  The reference reads ``runtime/Hi-C/sigma``, which
the current A-step no longer sets (quirk q2); ``intra_sigma`` is used when it is
absent.  PARITY UNPINNED: buildContactMap / sumCopies live in alabtools.
"""
from __future__ import annotations

import os

from ..contact import counts_to_probmatrix, evaluation_stats, haploid_contact_counts
from ..engine import ActdistEngine
from ..population import Population, ProbMatrix
from ._compat import Step, logger

eps = 0.05          # igm/steps/HicEvaluationStep.py:22


class HicEvaluationStep(Step):

    def _sigma(self):
        rt = self.cfg["runtime"].get("Hi-C", {})
        s = rt.get("sigma", rt.get("intra_sigma"))
        if s is None or s is False:
            s = rt.get("inter_sigma")
        if s is None or s is False:
            raise KeyError("runtime/Hi-C/sigma")
        return float(s)

    def name(self):
        s = 'HicEvaluationStep (sigma={:.2f}%, iter={:s})'
        return s.format(self._sigma() * 100.0, str(self.cfg.get('runtime/opt_iter', 'N/A')))

    def setup(self):
        self.out_dir = os.path.join(
            self.cfg.get('parameters/workdir'), 'evaluation', 'Hi-C',
            'sigma_{:.2f}.iter_{:s}'.format(self._sigma() * 100.0,
                                            str(self.cfg.get('runtime/opt_iter', 'NA'))))
        if not os.path.isdir(self.out_dir):
            os.makedirs(self.out_dir)
        self.argument_list = []

    @staticmethod
    def task(struct_id, cfg, tmp_dir):
        return

    def reduce(self):
        hic = self.cfg['restraints']['Hi-C']
        cr = float(hic.get('contact_range', 2.0)) * (1 + eps)      # :109
        pop = Population.from_hss(self.cfg.get('optimization/structure_output'))
        with ActdistEngine(pop, int(hic.get('gpu_device', 0))) as eng:
            counts = haploid_contact_counts(eng, cr, False)
        out = counts_to_probmatrix(counts, pop.nstruct, pop.chrom_hap(), clip=True)   # :111-112
        out.save_hcs(os.path.join(self.out_dir, 'out_matrix.hcs'))
        inp = ProbMatrix.from_hcs(self.cfg.get('restraints/Hi-C/input_matrix'))
        self.score, avg, avg_rel = evaluation_stats(inp, out, self._sigma())
        with open(os.path.join(self.out_dir, 'stats.txt'), 'w') as f:
            print("#score ave_differences ave_relative_differences", file=f)
            print(self.score, avg, avg_rel, file=f)
        logger.info('>>>  Average relative difference: {:6.3f}%  <<<'.format(self.score * 100))
