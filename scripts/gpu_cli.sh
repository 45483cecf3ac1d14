#!/bin/bash
set -u
mkdir -p gpurun_out /tmp/cli
export PYTHONUNBUFFERED=1
python - <<'PY'
import json, sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from helpers import load_case, build_product
meta, arr, weights = load_case("uvit_pose_vanilla")
json.dump(meta["cfg"], open("/tmp/cli/algo.json", "w"))
algo = build_product(meta["cfg"])
torch.save({"state_dict": {"diffusion_model.model." + k: v for k, v in weights.items()}, "pretrained_ema": True, "optimizer_states": []}, "/tmp/cli/m.ckpt")
vid = algo._unnormalize_x(torch.from_numpy(arr["xs"])).numpy()
np.savez("/tmp/cli/batch.npz", videos=np.concatenate([vid, vid], 0), conds=np.concatenate([arr["conds"], arr["conds"]], 0))
PY
python -m dfot_b200.experiments --config /tmp/cli/algo.json --ckpt /tmp/cli/m.ckpt --input /tmp/cli/batch.npz --output /tmp/cli/out1.npz --seed 3 2>&1 | tail -2
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 -m dfot_b200.experiments --config /tmp/cli/algo.json --ckpt /tmp/cli/m.ckpt --input /tmp/cli/batch.npz --output /tmp/cli/out2.npz --seed 3 --br 2 2>&1 | tail -2
fi
python - <<'PY'
import numpy as np, os
a = np.load("/tmp/cli/out1.npz"); print({k: a[k].shape for k in a.files}, float(np.abs(a["prediction"]).max()))
if os.path.exists("/tmp/cli/out2.npz"):
    b = np.load("/tmp/cli/out2.npz"); print("2-GPU branch-split vs 1-GPU max diff (different noise per shard expected: shapes only)", b["prediction"].shape)
PY
