"""ContinuousDiffusion — sampling-time drop-in for continuous_diffusion.py:94-167: identical to the discrete
sampler except that the backbone is fed ``precond_scale * logsnr[k]`` (fp32) instead of the integer level."""
import torch

from .discrete_diffusion import DiscreteDiffusion


class ContinuousDiffusion(DiscreteDiffusion):
    is_continuous = True

    def __init__(self, cfg, backbone_cfg, x_shape, max_tokens, external_cond_type, external_cond_num_classes,
                 external_cond_dim):
        super().__init__(cfg, backbone_cfg, x_shape, max_tokens, external_cond_type, external_cond_num_classes,
                         external_cond_dim)
        assert self.objective == "pred_v" and self.loss_weighting.strategy == "sigmoid", \
            "ContinuousDiffusion only supports 'pred_v' objective and 'sigmoid' loss weighting"
        self.precond_scale = self.cfg.precond_scale
        self.sigmoid_bias = self.cfg.loss_weighting.sigmoid_bias

    def model_input_levels(self, k: torch.Tensor) -> torch.Tensor:
        return self.precond_scale * self.logsnr[k]
