"""GPU: the C++ RetrievalModel plugin (REGISTER_MODEL(B200IVFPQ/B200FLAT)) against the reference's own models,
both obtained from the reference's reflector in one process (gamma_b200/plugin/plugin_parity_main.cc).
The binary is built here by `make -C oracle plugin` (needs the reference sources) and ships to the GPU box."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "gamma_b200", "plugin", "_build", "plugin_parity")


def test_plugin_matches_reference_models_through_the_reflector():
    if not os.path.exists(BIN):
        pytest.skip("plugin_parity not built (make -C oracle plugin needs /root/reference)")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_WAIT_POLICY="PASSIVE")
    p = subprocess.run([BIN], capture_output=True, text=True, timeout=600, env=env)
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, (p.returncode, p.stdout, p.stderr)
    assert p.returncode == 0, lines
    assert lines[-1]["plugin_parity"] == "PASS", lines
    cases = {l["case"]: l for l in lines if "case" in l}
    assert cases["flat_filter_delete_update"]["ids_identical"] == 1.0
    assert cases["ivfpq_rerank"]["ids_identical"] >= 0.995
    # docs deleted through the bitmap alone, host-side Update + CompactBucket mirrored list by list, Dump / Load in the
    # reference's file format in both directions, 8 concurrent searches
    for name in ("ivfpq_bitmap_only_delete", "flat_bitmap_only_delete", "ivfpq_rerank_after_update_and_compaction",
                 "ivfpq_adc_filter_after_compaction", "b200_loads_reference_dump", "reference_loads_b200_dump",
                 "loaded_b200_equals_live_b200", "concurrent_search_8_threads", "ivfflat", "ivfflat_filter_deleted",
                 "ivfflat_after_update"):
        assert cases[name]["ok"], cases[name]
    assert cases["loaded_b200_equals_live_b200"]["ids_identical"] == 1.0
    assert any(l.get("host_compacted_postings", 0) > 0 for l in lines)
