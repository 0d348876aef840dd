"""CPU: the batch form of the bi-predictive refinement (ora_mb_bipred, the block of mv-search.c:864-1034) against the chains of
FullPelBlockMotionBiPred / SubPelBlockSearchBiPred calls the unmodified reference made in the known-answer run."""
import numpy as np

from conftest import golden_trace
from trace_util import bipred_chains


def kat_chain_jobs(ob, tr):
    chains = bipred_chains(tr, refinements=3, subpel=2)      # bin/encoder.cfg: BiPredMERefinements 3, BiPredMESubPel 2
    jobs = np.zeros(len(chains), ob.BIPRED_JOB_DTYPE)
    for k, c in enumerate(chains):
        j = jobs[k]
        j["opix_x"], j["opix_y"], j["list"] = c["opix"][0], c["opix"][1], c["list"]
        j["pmv"], j["pmv_bi"] = c["pmv"], c["pmv_bi"]
        j["mv"] = (4 * c["mv_int"][0], 4 * c["mv_int"][1])   # any quarter-pel vector that rounds to the traced integer one
        j["lambda"] = c["lam"]
    return chains, jobs


def test_bipred_batch_job_reproduces_reference_chains(oracle_lib):
    ob = oracle_lib
    tr, cfgkw = golden_trace("kat")
    cfg = ob.make_cfg(**cfgkw)
    chains, jobs = kat_chain_jobs(ob, tr)
    assert len(chains) >= 20
    refs = {}
    n_checked = 0
    for c, job in zip(chains, jobs):
        assert c["search_range"] == 16
        lst = c["list"]
        ids = {lst: c["ref_l"][0], lst ^ 1: c["ref_l"][1]}
        for i in ids.values():
            if i not in refs:
                refs[i] = ob.Ref(tr.ref(i))
        cur = np.ascontiguousarray(tr.cur(c["pic"]))
        out = ob.mb_bipred(cfg, cur, refs[ids[0]], refs[ids[1]], jobs[n_checked:n_checked + 1], 3, 16, 2)[0]
        assert tuple(out["mv"]) == c["exp_mv"], (n_checked, out, c["exp_mv"])
        assert tuple(out["bimv"]) == c["exp_bimv"], (n_checked, out, c["exp_bimv"])
        assert int(out["cost"]) == c["exp_cost"], (n_checked, out, c["exp_cost"])
        assert int(out["iterlist"]) == lst ^ 1                # refinements = 3: the last pass ran on the other list
        n_checked += 1
    for r in refs.values():
        r.close()
