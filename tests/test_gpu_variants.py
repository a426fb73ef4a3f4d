"""Every switchable path of the sorter against the oracle: the text rounds (none, one, two), the generic and the DNA
classification, the staged and the register-blocked regroup kernels, the global radix sort instead of the in-tile sort
of a round, both lcp paths.  One subprocess per variant (tests/variant_worker.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "default": {},
    "no_text_round": {"ESA_TEXT_ROUNDS": "0"},
    "two_text_rounds": {"ESA_TEXT_ROUNDS": "2"},
    "generic_classify": {"ESA_NO_DNA_CLASSIFY": "1"},
    "regroup_staged": {"ESA_REGROUP_STAGED": "1"},
    "regroup_blocked": {"ESA_REGROUP_BLOCKED": "1"},
    "global_round_sort": {"ESA_NO_LOCAL_SORT": "1"},
    "plcp_path": {"ESA_LCP_MODE": "plcp"},
    "plcp_no_text_round": {"ESA_LCP_MODE": "plcp", "ESA_TEXT_ROUNDS": "0"},
    # hybrid lcp: slots split at a depth of at least ESA_LCP_DEEP symbols take the text-order value (default 1024, which
    # the small inputs rarely reach: force it down); also on a sharded context, where the deep values travel between ranks
    "dna_emitted_first_pass": {"ESA_NO_TEXT_PASS": "1"},  # initial DNA sort from an emitted (key, position) list
    "dna_emitted_first_pass_sharded": {"ESA_NO_TEXT_PASS": "1", "VARIANT_RANKS": "3"},
    "protein_byte_text": {"ESA_NO_PACK5": "1"},  # alphabets of 5..30 letters without the 5-bit records
    "lcp_direct_only": {"ESA_LCP_MODE": "direct"},
    "lcp_hybrid_deep16": {"ESA_LCP_DEEP": "16"},
    "lcp_hybrid_deep16_sharded": {"ESA_LCP_DEEP": "16", "VARIANT_RANKS": "3"},
    "lcp_hybrid_deep40_no_text_round_sharded": {"ESA_LCP_DEEP": "40", "ESA_TEXT_ROUNDS": "0", "VARIANT_RANKS": "2"},
    "lcp_all_text_order_sharded": {"ESA_LCP_MODE": "plcp", "VARIANT_RANKS": "4"},
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_sorter_variant_matches_oracle(name):
    env = dict(os.environ)
    for k in ("ESA_TEXT_ROUNDS", "ESA_NO_DNA_CLASSIFY", "ESA_REGROUP_STAGED", "ESA_REGROUP_BLOCKED", "ESA_NO_LOCAL_SORT",
              "ESA_LCP_MODE", "ESA_LCP_DEEP", "VARIANT_RANKS", "ESA_NO_PACK5", "ESA_NO_TEXT_PASS"):
        env.pop(k, None)
    env.update(VARIANTS[name])
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_worker.py")], capture_output=True,
                        text=True, timeout=900, env=env, cwd=ROOT)
    assert pr.returncode == 0 and "VARIANT-OK" in pr.stdout, pr.stdout[-2000:] + pr.stderr[-3000:]
