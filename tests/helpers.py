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
