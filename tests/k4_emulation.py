"""TEST INFRASTRUCTURE: a torch/CPU restatement of the *contract* of dfot_sampler_step_hg
(include/dfot_b200.h), used to check the host-side window planner on machines without a GPU and as the
reference the CUDA kernel is compared against in the -m gpu tests.  Never imported by the product."""
import numpy as np
import torch

UPDATE_DTYPE = np.dtype([("a", "<f4"), ("b", "<f4"), ("sigma", "<f4"), ("w", "<f4"), ("clip", "<f4"),
                         ("generate", "<i4")])
PREPARE_DTYPE = np.dtype([("mode", "<i4"), ("noise_row", "<i4"), ("qa", "<f4"), ("qb", "<f4")])


def _table(t, dtype, rows, T):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.frombuffer(t.tobytes(), dtype=dtype).reshape(rows, T)


def emulate(x, model_out, model_in_next, upd, prep, noise_ddim, noise_hist, noise_excl, B, nfe, T, max_noise_row=None):
    """In-place on x (f32 [B,T,...]) and model_in_next, like the kernel."""
    F = x[0, 0].numel()
    xs = x.reshape(B, T, F)
    upd = _table(upd, UPDATE_DTYPE, B * nfe, T)
    prep = _table(prep, PREPARE_DTYPE, B * nfe, T)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
    if model_out is not None:
        out = model_out.float().reshape(B, nfe, T, F)
        a, b, sg, w, clip = (f32(upd[k]).reshape(B, nfe, T, 1) for k in ("a", "b", "sigma", "w", "clip"))
        gen = torch.from_numpy(upd["generate"].reshape(B, nfe, T)[:, 0].copy()).bool()
        o = torch.where(clip > 0, torch.maximum(torch.minimum(out, clip), -clip), out)
        v = torch.where(b == 0, a * xs[:, None], a * xs[:, None] + b * o)     # b == 0: kept frame, out is not read
        if noise_ddim is not None:
            v = v + sg * noise_ddim.float().reshape(B, nfe, T, F)
        comp = (w * v).sum(1)
        xs.copy_(torch.where(gen[..., None], comp, xs))
    if model_in_next is not None:
        mode = torch.from_numpy(prep["mode"].reshape(B, nfe, T).copy())
        row = prep["noise_row"].reshape(B, nfe, T)
        qa, qb = f32(prep["qa"]).reshape(B, nfe, T, 1), f32(prep["qb"]).reshape(B, nfe, T, 1)
        res = xs[:, None].expand(B, nfe, T, F).clone()
        if noise_hist is not None:
            nh = noise_hist.float().reshape(-1, T, F)
            gathered = nh[torch.from_numpy(row.astype(np.int64)), torch.arange(T)[None, None, :]]
            res = torch.where((mode == 1)[..., None], qa * xs[:, None] + qb * gathered, res)
        if noise_excl is not None:
            res = torch.where((mode == 2)[..., None], noise_excl.float().reshape(B, nfe, T, F), res)
        model_in_next.copy_(res.reshape(model_in_next.shape).to(model_in_next.dtype))
