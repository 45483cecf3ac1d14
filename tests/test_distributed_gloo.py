"""World-size-2 gloo tests (CPU) of the multi-GPU sharding logic: sample sharding and history-guidance branch
sharding must reproduce the single-process rollout exactly.  Kernels are replaced by the K4 contract emulation and
the oracle backbone (host logic is what is under test here; the CUDA path is covered by the -m gpu tests)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(case):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import k4_emulation
    from dfot_b200 import ops
    from helpers import NoiseBank, build_oracle, build_product, load_case
    meta, arr, weights = load_case(case)
    cfg = meta["cfg"]
    algo = build_product(cfg)
    algo.model_in_dtype = torch.float32
    if cfg["backbone"]["name"] == "u_vit3d_pose":
        # the product's own U-ViT3DPose host code (pose cache, row maps under branch sharding) on emulated kernels
        import ops_emulation
        sd = {"diffusion_model.model." + k: v for k, v in weights.items()}
        sd["data_mean"], sd["data_std"] = algo.data_mean, algo.data_std
        algo.load_state_dict(sd, strict=True)
        ops_emulation.install_raw()
    else:
        _, backbone = build_oracle(cfg, weights)

        class OracleBackbone(torch.nn.Module):
            def forward(self, x, k, c=None, cm=None, out_dtype=None):
                return backbone(x, k, c, cm)

        algo.diffusion_model.model = OracleBackbone()
    ops.sampler_step_hg = k4_emulation.emulate
    xs = torch.from_numpy(arr["xs"])
    conds = torch.from_numpy(arr["conds"]) if "conds" in arr else None
    return algo, cfg, xs, conds, NoiseBank


def _worker(rank, world, port, case, br, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        algo, cfg, xs, conds, NoiseBank = _build(case)
        from dfot_b200 import distributed as D
        algo.mesh = D.build_mesh(br=br)
        # per-sample noise streams so that a sample's result does not depend on which shard it lands in; with fewer
        # samples than shards (interpolation chunk sharding) every rank replays the same stream
        sl = D.shard_batch(xs.shape[0], algo.mesh.dp, algo.mesh.dp_index)
        bank = NoiseBank(100 + (sl.start if xs.shape[0] >= algo.mesh.dp else 0))
        algo.diffusion_model.noise_source = bank          # (forkable: interpolation rounds advance in lockstep)
        out = algo.sample_sharded(xs, conds, cfg["context_frames"])
        np.save(out_path + f".rows{rank}.npy", np.array([algo.nfe_rows]))
        if rank == 0:
            np.save(out_path, out.numpy())
    finally:
        dist.destroy_process_group()


def _single(case, shard_starts):
    sys.path.insert(0, ROOT)
    from dfot_b200 import ops
    real_ops = dict(vars(ops))
    try:
        algo, cfg, xs, conds, NoiseBank = _build(case)   # patches ops.sampler_step_hg in this process
        outs = []
        for a, b in zip(shard_starts[:-1], shard_starts[1:]):
            if a == b:
                continue
            bank = NoiseBank(100 + a)
            algo.diffusion_model.noise_source = lambda shape, device, bank=bank: bank.randn(shape)
            outs.append(algo._predict_videos(xs[a:b].contiguous(), cfg["context_frames"],
                                             None if conds is None else conds[a:b]))
        _single.nfe_rows = algo.nfe_rows
        return torch.cat(outs, 0).numpy()
    finally:
        for k, v in real_ops.items():
            setattr(ops, k, v)


@pytest.mark.parametrize("case,br", [("vanilla", 1), ("vanilla", 2), ("continuous_action", 2), ("temporal", 1),
                                     ("uvit_pose_vanilla", 2), ("label_vanilla", 1), ("label_vanilla", 2),
                                     ("refine_conditional", 1)])
def test_world2_matches_single_process(case, br, tmp_path):
    world = 2
    port = 29500 + (os.getpid() + hash((case, br))) % 2000
    out_path = str(tmp_path / "out.npy")
    mp.spawn(_worker, args=(world, port, case, br, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    dp = world // br
    import json
    with open(os.path.join(ROOT, "tests", "golden", f"case_{case}.json")) as f:
        batch = json.load(f)["batch"]
    sys.path.insert(0, ROOT)
    from dfot_b200 import distributed as D
    starts = [D.shard_batch(batch, dp, d).start for d in range(dp)] + [batch]
    want = _single(case, starts)
    assert got.shape == want.shape
    if br == 1:
        assert np.array_equal(got, want)          # sample sharding: bit-exact
    else:                                         # branch sharding changes the CPU BLAS batch shape of the checker
        assert np.abs(got - want).max() <= (2e-2 if "uvit" in case else 1e-5)   # (bf16-emulated kernels for U-ViT)


@pytest.mark.parametrize("case", ["keyframes_interp", "uvit_pose_stabilized_interp"])
def test_world2_interpolation_chunks_are_sharded(case, tmp_path):
    """One sample, two ranks: the sampler state is replicated, the forward rows of the keyframe windows and of every
    interpolation round (all chunk batches in lockstep) are dealt over the ranks, and the rollout must equal the
    single-process one — bit for bit with the DiT oracle backbone, to the bf16 emulation's noise floor with the U-ViT
    (the CPU BLAS blocks differently for another batch shape)."""
    world = 2
    port = 29500 + (os.getpid() + hash((case, "chunks"))) % 2000
    out_path = str(tmp_path / "out.npy")
    mp.spawn(_worker, args=(world, port, case, 1, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    want = _single(case, [0, 1])
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= (2e-2 if "uvit" in case else 1e-5)
    rows = [int(np.load(out_path + f".rows{r}.npy")[0]) for r in range(world)]
    assert sum(rows) == _single.nfe_rows and max(rows) <= (_single.nfe_rows + 1) // 2 + 8   # every row forwarded once


def _seeded_worker(rank, world, port, case, seed, duplicate, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        algo, cfg, xs, conds, _ = _build(case)
        from dfot_b200 import distributed as D
        algo.mesh = D.build_mesh(br=1)
        if duplicate:                      # two copies of sample 0: only the noise can tell the shards apart
            xs = xs[:1].repeat(2, *[1] * (xs.dim() - 1))
            conds = None if conds is None else conds[:1].repeat(2, *[1] * (conds.dim() - 1))
        torch.manual_seed(seed)            # what SamplingExperiment(manual_seed=seed) does: the SAME seed on every rank
        out = algo.sample_sharded(xs, conds, cfg["context_frames"])
        np.save(out_path + f".{rank}.npy", out.numpy())
    finally:
        dist.destroy_process_group()


def test_world2_seeded_single_sample_rollout_is_coherent(tmp_path):
    """with a user seed and fewer samples than dp shards, keyframe windows are replicated and
    interpolation chunks are dealt over the shards — every rank must draw ONE noise stream, end with the same video, and
    that video must be the single-process rollout of the same seed."""
    case, seed = "keyframes_interp", 77
    out_path = str(tmp_path / "seeded")
    mp.spawn(_seeded_worker, args=(2, 29500 + (os.getpid() + 11) % 2000, case, seed, False, out_path), nprocs=2, join=True)
    got = [np.load(out_path + f".{r}.npy") for r in range(2)]
    assert np.array_equal(got[0], got[1])
    sys.path.insert(0, ROOT)
    from dfot_b200 import ops
    real_ops = dict(vars(ops))
    try:
        algo, cfg, xs, conds, _ = _build(case)
        torch.manual_seed(seed)
        want = algo._predict_videos(xs, cfg["context_frames"], conds).numpy()
    finally:
        for k, v in real_ops.items():
            setattr(ops, k, v)
    assert np.abs(got[0] - want).max() <= 1e-5      # (the CPU checker's BLAS blocks differently for other row counts)


def test_world2_seeded_sample_shards_draw_different_noise(tmp_path):
    """Same seed on every rank, samples sharded: each dp shard derives its own stream (two copies of one input must not
    come out identical), every rank ends with the same gathered batch, and a second run reproduces it."""
    case, seed = "vanilla", 5
    runs = []
    for attempt in range(2):
        out_path = str(tmp_path / f"dup{attempt}")
        mp.spawn(_seeded_worker, args=(2, 29500 + (os.getpid() + 23 + attempt) % 2000, case, seed, True, out_path),
                 nprocs=2, join=True)
        got = [np.load(out_path + f".{r}.npy") for r in range(2)]
        assert np.array_equal(got[0], got[1])
        runs.append(got[0])
    assert np.array_equal(runs[0], runs[1])
    assert np.abs(runs[0][0] - runs[0][1]).max() > 1e-3


# ------------------------------------------------------------------------------------------------ sharded VAE decode
def _vae_for_gloo():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ops_emulation
    from dfot_b200.algorithms.vae import VideoVAE
    from oracle.video_vae import decoder_param_shapes, seeded_weights
    vae = VideoVAE(hidden_size=32, z_channels=4, embed_dim=4, hidden_size_mult=(1, 2, 2, 2))
    vae.load_state_dict(seeded_weights(decoder_param_shapes(32, 4, 4, (1, 2, 2, 2)), 11))
    restore = ops_emulation.install_raw()
    z = torch.randn((5, 4, 2, 4, 4), generator=torch.Generator().manual_seed(3))      # 5 clips: uneven over 2 and 3 ranks
    return vae, z, restore


def _decode_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vae, z, _ = _vae_for_gloo()
        from dfot_b200 import distributed as D
        calls = []

        def decode(lat):
            calls.append(lat.shape[0])
            return vae.decode(lat, 5)
        out = D.decode_sharded(decode, z)
        np.save(out_path + f".{rank}.npy", out.numpy())
        np.save(out_path + f".calls{rank}.npy", np.array(calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_vae_decode_sharded_over_ranks_equals_single_process(world, tmp_path):
    """`decode_sharded` (the decode after the final sample gather): every rank decodes ceil(5 / world) clips, one
    all_gather, and every rank ends with the single-process result bit for bit (kernels: CPU contract emulations; the
    single process decodes the same chunks, since the CPU convolution's blocking — hence its rounding — depends on the
    batch size)."""
    from dfot_b200 import ops
    real_ops = dict(vars(ops))
    per = -(-5 // world)
    try:
        vae, z, _ = _vae_for_gloo()
        chunks = [torch.arange(r * per, (r + 1) * per).clamp(max=4) for r in range(world)]
        ref = torch.cat([vae.decode(z[i], 5) for i in chunks], 0)[:5].numpy()
    finally:
        for k, v in real_ops.items():
            setattr(ops, k, v)
    out_path = str(tmp_path / "dec")
    port = 29640 + world
    mp.spawn(_decode_worker, args=(world, port, out_path), nprocs=world, join=True)
    for r in range(world):
        got = np.load(out_path + f".{r}.npy")
        assert got.shape == ref.shape == (5, 3, 5, 32, 32)
        assert np.array_equal(got, ref)
        assert list(np.load(out_path + f".calls{r}.npy")) == [per]
