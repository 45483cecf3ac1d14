"""Reading checkpoint files of the reference (`.ckpt`): one place that decides how much of pickle is trusted."""
from typing import Dict, Optional

import torch


def load_checkpoint_file(path: str, allow_pickle: Optional[bool] = None) -> Dict:
    """`torch.load` of a `.ckpt` with the tensor-only unpickler first: the checkpoints this path targets (EMA-only releases,
    Lightning / Accelerate state dicts) need tensors, lists, dicts and scalars only, and a downloaded file must not be able
    to run code.  A checkpoint that pickles other objects (e.g. hyper-parameter containers) loads only after an explicit
    opt-in: `allow_pickle=True` or DFOT_ALLOW_PICKLE_CKPT=1."""
    import os
    import pickle
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except (pickle.UnpicklingError, RuntimeError) as e:
        if allow_pickle is None:
            allow_pickle = os.environ.get("DFOT_ALLOW_PICKLE_CKPT") == "1"
        if not allow_pickle:
            raise RuntimeError(
                f"{path} holds pickled objects beyond tensors and plain containers ({str(e).splitlines()[0]}); loading it "
                "would execute code from the file.  If you trust its source, set DFOT_ALLOW_PICKLE_CKPT=1 (or pass "
                "allow_pickle=True) to load it with the full unpickler.") from e
        return torch.load(path, map_location="cpu", weights_only=False)
