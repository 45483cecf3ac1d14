"""Checkpoint ingestion: reference `.ckpt` / `.safetensors` files load through the same key
rules as the reference's on_load_checkpoint (EMA swap, torch.compile prefixes, key filtering, strictness)."""
import pytest
import torch

from helpers import build_product, load_case


def _algo_and_weights(name="vanilla"):
    meta, _, weights = load_case(name)
    return build_product(meta["cfg"]), {"diffusion_model.model." + k: v for k, v in weights.items()}, meta["cfg"]


def _model_state(algo):
    return {k: v.clone() for k, v in algo.state_dict().items() if k.startswith("diffusion_model.model.")}


@pytest.mark.parametrize("name", ["vanilla", "uvit_pose_vanilla"])
def test_lightning_ckpt_with_ema_and_compile_prefix(name, tmp_path):
    algo, sd, _ = _algo_and_weights(name)
    raw = {k.replace("diffusion_model.", "diffusion_model._orig_mod."): torch.zeros_like(v) for k, v in sd.items()}
    raw["data_mean"] = torch.full_like(algo.data_mean, 7.0)               # ignored: not a model key
    params = [sd["diffusion_model." + k] for k, _ in algo.diffusion_model.named_parameters()]
    path = str(tmp_path / "full.ckpt")
    torch.save({"state_dict": raw, "optimizer_states": [{"ema": params}], "epoch": 3}, path)
    algo.load_checkpoint(path)
    got = _model_state(algo)
    for k, _ in algo.diffusion_model.named_parameters():
        assert torch.equal(got["diffusion_model." + k], sd["diffusion_model." + k]), k      # EMA replaced the zeros
    assert "data_mean" in algo.ckpt_ignored_keys and not algo.ckpt_missing_keys
    assert not torch.equal(algo.data_mean, raw["data_mean"])


def test_ema_only_release_ckpt_and_safetensors(tmp_path):
    from safetensors.torch import save_file
    algo, sd, _ = _algo_and_weights()
    p1 = str(tmp_path / "release.ckpt")
    torch.save({"state_dict": sd, "pretrained_ema": True, "optimizer_states": []}, p1)
    algo.load_checkpoint(p1)
    assert all(torch.equal(v, sd[k]) for k, v in _model_state(algo).items())
    algo2, _, _ = _algo_and_weights()
    p2 = str(tmp_path / "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, p2)
    algo2.load_checkpoint(p2)
    assert all(torch.equal(v, sd[k]) for k, v in _model_state(algo2).items())


def test_missing_keys_follow_strict_flag(tmp_path):
    algo, sd, cfg = _algo_and_weights()
    dropped = next(k for k in sd if k.endswith("final_layer.linear.weight"))
    partial = {k: v for k, v in sd.items() if k != dropped}
    path = str(tmp_path / "partial.ckpt")
    torch.save({"state_dict": partial, "pretrained_ema": True, "optimizer_states": []}, path)
    with pytest.raises(ValueError, match="not found in the checkpoint"):
        algo.load_checkpoint(path)
    cfg["checkpoint"]["strict"] = False
    lenient = build_product(cfg)
    before = lenient.state_dict()[dropped].clone()
    lenient.load_checkpoint(path)
    assert lenient.ckpt_missing_keys == [dropped] and torch.equal(lenient.state_dict()[dropped], before)
    saved = {"state_dict": dict(lenient.state_dict())}
    lenient.on_save_checkpoint(saved)
    assert all(k.startswith("diffusion_model.model") for k in saved["state_dict"])


def test_uvit3d_pose_parameter_order_matches_the_reference():
    """The EMA list of a Lightning checkpoint is zipped with `named_parameters()` (base_pytorch_video_algo.py:1130-1144 of the
    reference): the product's order must be the reference's, or same-shaped tensors would be swapped silently.  The golden
    weight file was written from the executed reference's `state_dict()` (oracle/make_goldens.py), whose key order is its
    named_parameters() order with the persistent buffers in between."""
    import numpy as np
    import os
    from helpers import GOLDEN
    for case, wfile in [("uvit_pose_vanilla", "weights_uvit_pose.npz"), ("vanilla", "weights_plain.npz"),
                        ("continuous_action", "weights_action.npz"), ("label_vanilla", "weights_label.npz"),
                        ("factorized_vanilla", "weights_factorized.npz"), ("learned_1d_vanilla", "weights_learned_1d.npz"),
                        ("matrix_factorized_bias", "weights_matrix_factorized_bias.npz"),
                        ("matrix_self_factorized", "weights_matrix_self.npz"),
                        ("matrix_cross_full", "weights_matrix_cross.npz"),
                        ("matrix_bias_cols2", "weights_matrix_bias_cols2.npz")]:
        algo, _, _ = _algo_and_weights(case)
        ref_keys = list(np.load(os.path.join(GOLDEN, wfile)).files)
        model = algo.diffusion_model.model
        assert list(model.state_dict().keys()) == ref_keys, case
        own = [k for k, _ in model.named_parameters()]
        params = set(own)
        assert own == [k for k in ref_keys if k in params], case


def test_pickled_objects_need_an_opt_in(tmp_path, monkeypatch):
    """A `.ckpt` is loaded with the tensor-only unpickler; arbitrary pickled objects load only after an explicit opt-in."""
    import argparse
    algo, sd, _ = _algo_and_weights()
    path = str(tmp_path / "objects.ckpt")
    torch.save({"state_dict": sd, "pretrained_ema": True, "optimizer_states": [], "hparams": argparse.Namespace(a=1)}, path)
    monkeypatch.delenv("DFOT_ALLOW_PICKLE_CKPT", raising=False)
    with pytest.raises(RuntimeError, match="DFOT_ALLOW_PICKLE_CKPT"):
        algo.load_checkpoint(path)
    monkeypatch.setenv("DFOT_ALLOW_PICKLE_CKPT", "1")
    algo.load_checkpoint(path)
    assert all(torch.equal(v, sd[k]) for k, v in _model_state(algo).items())
