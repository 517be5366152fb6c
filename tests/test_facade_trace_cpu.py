"""CPU checks of the recorded reference-facade trace (tests/golden/facade_trace.json).

The trace is what /root/reference/gaussian_renderer/__init__.py -- executed unmodified with a recording rasterizer --
passes to and expects from `diff_gaussian_rasterization_depth_alpha` (tests/golden/make_facade_trace.py).  Here:
 * the committed trace is internally consistent with this repo's host-side mirror (keyword sets, field order);
 * whenever /root/reference is present (authoring container) the trace is regenerated from the reference and must
   equal the committed one, so the fixture cannot drift from the facade it claims to record.
The GPU replay of every call is tests/test_gpu_parity.py::test_reference_facade_trace_replay.
"""
import json
import os
import subprocess
import sys

import pytest

import helpers as h

HERE = os.path.dirname(os.path.abspath(__file__))
TRACE = os.path.join(HERE, "golden", "facade_trace.json")


def _trace():
    with open(TRACE) as f:
        return json.load(f)


def test_trace_keyword_sets_match_the_host_mirror():
    t = _trace()
    fields = list(h.dgr.GaussianRasterizationSettings._fields)
    import inspect
    call_params = [p for p in inspect.signature(h.dgr.GaussianRasterizer.forward).parameters if p != "self"]
    assert len(t["scenarios"]) >= 8
    for sc in t["scenarios"]:
        assert sc["settings_keys"] == fields, sc["name"]                  # same 12 names, same order
        assert sorted(sc["call_keys"]) == sorted(call_params), sc["name"]  # the 8 call kwargs
        c = sc["call"]
        assert (c["shs"] is None) != (c["colors_precomp"] is None)
        assert (c["cov3D_precomp"] is None) != (c["scales"] is None and c["rotations"] is None)
        for k, m in c.items():
            if m is not None:
                assert m["dtype"] == "float32" and m["contiguous"], (sc["name"], k)
    r = t["scenarios"][0]["returns"]
    assert r["keys"] == ["render", "viewspace_points", "visibility_filter", "radii", "3dposition", "depth", "alpha"]
    assert r["visibility_filter_is_radii_gt_0"] and r["position_is_call_means3D"]


def test_trace_covers_every_facade_branch():
    names = {sc["name"]: sc for sc in _trace()["scenarios"]}
    assert names["render_convert_SHs_python"]["call"]["shs"] is None                 # gaussian_renderer/__init__.py:80-85
    assert names["render_compute_cov3D_python"]["call"]["scales"] is None            # :69-70
    assert names["render_compute_cov3D_python"]["settings_scalars"]["scale_modifier"] == 1.3
    assert names["render_vis_mask_avatar"]["call"]["means3D"]["shape"][0] < _trace()["N"]   # :92-100
    assert names["render_override_color_masked"]["call"]["colors_precomp"]["requires_grad"] is False
    assert names["doll_override_shs_masked"]["fn"] == "doll_render"                  # :189-192
    assert names["doll_default"]["returns"]["order"] == ["rendered_image", "depth", "alpha"]   # :221
    assert all(not m["requires_grad"] for m in names["doll_default"]["call"].values() if m)    # inference.py:462


@pytest.mark.skipif(not os.path.isdir("/root/reference/gaussian_renderer"), reason="reference tree not present")
def test_committed_trace_equals_a_fresh_run_of_the_reference_facade(tmp_path):
    env = dict(os.environ, GG_FACADE_TRACE_OUT=str(tmp_path))
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_facade_trace.py")], env=env,
                          stdout=subprocess.DEVNULL)
    with open(tmp_path / "facade_trace.json") as f:
        fresh = json.load(f)
    assert fresh == _trace()
