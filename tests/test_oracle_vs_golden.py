"""Pin the oracle against the reference-generated goldens (CPU; no GPU needed).

Integers bit-exact; fp32 tensors <= 1e-5 (the oracle restates the same fp32 math in a
different op order, e.g. permute+reshape instead of einops)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import history_guidance as hg
from oracle import schedule
from oracle.sampler import interpolation_plan
from helpers import GOLDEN, build_oracle, case_names, load_case

with open(os.path.join(GOLDEN, "integers.json")) as f:
    INTS = json.load(f)


def test_scheduling_matrices_bit_exact():
    assert len(INTS["scheduling_matrices"]) >= 20
    for rec in INTS["scheduling_matrices"]:
        m = schedule.scheduling_matrix(rec["kind"], rec["horizon"], rec["padding"], 1000, rec["steps"])
        assert m.dtype == torch.int64
        assert m.tolist() == rec["matrix"], rec["kind"]


def test_ddim_levels_bit_exact():
    for rec in INTS["ddim_levels"]:
        lv = schedule.ddim_idx_to_noise_level(torch.arange(rec["steps"] + 1), 1000, rec["steps"])
        assert lv.tolist() == rec["levels"]


def test_known_answers_from_survey():
    # SURVEY.md appendix A.1/A.2 (probed from the reference)
    assert schedule.ddim_idx_to_noise_level(torch.arange(11), 1000, 10).tolist() == [-1, 99, 199, 299, 399, 499, 599, 699, 799, 899, 999]
    m = schedule.scheduling_matrix("autoregressive", 3, 2, 1000, 4)
    assert m.tolist() == [[999] * 5, [749, 999, 999, 999, 999], [499, 749, 999, 999, 999], [249, 499, 749, 999, 999],
                          [-1, 249, 499, 999, 999], [-1, -1, 249, 999, 999], [-1, -1, -1, 999, 999]]


def test_hg_branch_tables_bit_exact():
    n_full = 0
    for rec in INTS["hg_branch_tables"]:
        scheme = hg.scheme_from_config(rec["scheme"], 1000)
        mask = torch.tensor(rec["mask"])
        if "error" in rec:
            with pytest.raises((AssertionError, IndexError)):
                hg.branch_table(scheme, mask)
            continue
        assert scheme.is_simple == (rec["manager"] == "SimpleHistoryGuidanceManager"), rec
        if scheme.is_simple:
            assert rec["nfe"] == (1 if scheme.hist_weights[0] == 1 else 2)
            continue
        tab = hg.branch_table(scheme, mask)
        n_full += 1
        assert tab.nfe == rec["nfe"]
        assert tab.hist_indices.tolist() == rec["hist_indices"]
        assert tab.gen_indices.tolist() == rec["gen_indices"]
        assert tab.gen_mask.long().tolist() == rec["gen_mask"]
        assert tab.hist_noise_levels.tolist() == rec["hist_noise_levels"], rec
        assert tab.cond_mask.long().tolist() == rec["cond_mask"]
        assert tab.weights.tolist() == rec["weights"]
    assert n_full >= 40


def test_interpolation_plans_bit_exact():
    for rec in INTS["interpolation_calls"]:
        T = rec["n_frames"]
        known = torch.zeros(T, dtype=torch.bool)
        known[rec["keyframes"]] = True
        plan = interpolation_plan(known, rec["max_tokens"])
        # the reference was intercepted at _sample_sequence: one call per round (max_batch_size=None),
        # mask rows = padded per-chunk "known" flags
        assert len(plan) == len(rec["calls"])
        k = known.clone()
        for round_, call in zip(plan, rec["calls"]):
            rows = []
            for frames in round_:
                r = k[frames].long().tolist()
                r = r + [r[-1]] * (rec["max_tokens"] - len(r))
                rows.append(r)
            assert rows == call["mask"]
            for frames in round_:
                k[frames] = True


def test_diffusion_buffers():
    from oracle.cases import algorithm_cfg, continuous_overrides
    gold = np.load(os.path.join(GOLDEN, "schedules.npz"))
    for tag, over in [("cosine", {}), ("continuous", continuous_overrides()),
                      ("sigmoid_zt", {"diffusion.beta_schedule": "sigmoid", "diffusion.schedule_fn_kwargs": {}}),
                      ("cosine_shift", {"diffusion.schedule_fn_kwargs": dict(shift=0.5)})]:
        buf = schedule.diffusion_buffers(algorithm_cfg(**over)["diffusion"])
        for name in ["alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "logsnr"]:
            key = f"{tag}.{name}"
            if key in gold:
                assert np.array_equal(buf[name].numpy(), gold[key]), key


@pytest.mark.parametrize("name", case_names())
def test_rollout_matches_reference(name):
    meta, arr, weights = load_case(name)
    cfg = meta["cfg"]
    oracle, _ = build_oracle(cfg, weights)
    oracle.trace = []
    xs = torch.from_numpy(arr["xs"])
    conds = torch.from_numpy(arr["conds"]) if "conds" in arr else None
    torch.manual_seed(meta["sampling_seed"])
    out = oracle.predict_videos(xs.clone(), cfg["context_frames"], conds)
    assert len(oracle.trace) == int(arr["n_steps"])
    for i, t in enumerate(oracle.trace):
        p = f"step{i:03d}."
        assert np.array_equal(t["levels_from"].numpy(), arr[p + "levels_from"]), (name, i)
        assert np.array_equal(t["levels_to"].numpy(), arr[p + "levels_to"]), (name, i)
        if t["cond_mask"] is None:
            assert p + "cond_mask" not in arr
        else:
            assert np.array_equal(t["cond_mask"].numpy(), arr[p + "cond_mask"])
        for k in ["model_in", "model_out", "step_out"]:
            err = np.abs(t[k].numpy() - arr[p + k]).max()
            assert err <= 1e-5, (name, i, k, err)
    err = np.abs(out.numpy() - arr["prediction"]).max()
    assert err <= 1e-5, (name, err)


def test_oracle_matrix_combinations_against_the_executed_reference():
    """The matrix-attention combinations the product is tested on beyond the goldens (oracle/cases.py MATRIX_COMBOS) pin the
    oracle LIVE against the executed reference backbone wherever the reference is present (/root/reference here, its unmodified
    copy oracle/_ref/reference elsewhere) — in a subprocess, because oracle/ref_shim.py installs stand-in modules."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probe = subprocess.run([sys.executable, "-c", "from oracle import ref_shim; print(ref_shim.available())"], cwd=root,
                           capture_output=True, text=True, timeout=300)
    if probe.stdout.strip() != "True":
        pytest.skip("reference not present (run oracle/build_ref.py in the authoring container)")
    run = subprocess.run([sys.executable, "-m", "oracle.check_matrix_combos"], cwd=root, capture_output=True, text=True,
                         timeout=600)
    assert run.returncode == 0 and "OK:" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]


@pytest.mark.parametrize("name", ["k600", "dmlab", "re10k"])
def test_oracle_fullsize_rollout_against_the_executed_reference(name):
    """oracle/check_fullsize.py: one DDIM step of the FULL-SIZE benchmarked models (bench.k600_cfg / bench.dmlab_cfg / bench.re10k_cfg) through
    the executed reference's `_predict_videos` and through the oracle, on the product's own state dict loaded strictly into
    the reference.  K600 takes ~45 s of CPU; RE10K (~2 min on 4 cores) runs when DFOT_SLOW_TESTS=1 — its last output is
    committed as profiles/r02_fullsize_oracle_vs_reference.txt."""
    import subprocess
    import sys
    if name == "re10k" and os.environ.get("DFOT_SLOW_TESTS") != "1":
        pytest.skip("set DFOT_SLOW_TESTS=1 (about 2 minutes of CPU)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probe = subprocess.run([sys.executable, "-c", "from oracle import ref_shim; print(ref_shim.available())"], cwd=root,
                           capture_output=True, text=True, timeout=300)
    if probe.stdout.strip() != "True":
        pytest.skip("reference not present (run oracle/build_ref.py in the authoring container)")
    run = subprocess.run([sys.executable, "-m", "oracle.check_fullsize", name], cwd=root, capture_output=True, text=True,
                         timeout=1800)
    assert run.returncode == 0 and "OK" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]
