"""CPU: the oracle's restatement of p7_DomainDecoding + the region trigger (oracle/p7_oracle.c: p7o_domain_decoding)
against the unmodified reference (oracle/_ref: p7_DomainDecoding, impl_sse/decoding.c:166-200, and the region walk of
p7_domaindef_ByPosteriorHeuristics, p7_domaindef.c:412-487) on the reference's own Forward/Backward specials."""
import os
import numpy as np
import pytest

from conftest import bits, golden_set, GOLDEN
from biobench2_b200 import seqdb
from oracle import port, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libp7ref.so not built")


@pytest.mark.parametrize("name", ["globins4", "fn3", "Pkinase"])
def test_decoding_bit_exact_and_trigger_is_necessary_for_a_region(name):
    g = golden_set(name)
    rp = ref.RefProfile.open(os.path.join(GOLDEN, "%s.hmm" % name))
    o = port.OracleProfile(g.oprofile.view(), g.oprofile)
    db = g.db
    ntrig = nreg = 0
    for s in range(0, db.nseq, max(1, db.nseq // 120)):
        L = int(db.lengths[s])
        if L == 0:
            continue
        dsq = db.residues[db.offsets[s] - 1: db.offsets[s] + L + 1]
        fb = rp.fwdbck(dsq, L)
        if fb["status"] != 0:
            continue
        a = rp.domain_decoding(L, fb["fwd_xmx"], fb["bck_xmx"], fb["bck_own_scales"])
        b = o.domain_decoding(L, fb["fwd_xmx"], fb["bck_xmx"], fb["bck_own_scales"])
        assert a["status"] == b["status"]
        for k in ("btot", "etot", "mocc"):
            assert np.array_equal(bits(a[k]), bits(b[k])), (name, s, k)
        assert a["triggered"] == b["triggered"]
        assert bits(b["mocc_max"]) == bits(a["mocc"][1:].max())
        if not a["triggered"]:
            assert a["nregions"] == 0            # no trigger => no region (p7_domaindef.c:419-421)
        ntrig += a["triggered"]; nreg += a["nregions"] > 0
    assert ntrig > 0 and nreg > 0                 # the set holds real homologs as well as random targets
    o.close(); rp.close()
