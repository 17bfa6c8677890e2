"""Stub of diffusers.models.modeling_utils.ModelMixin."""
import torch


class ModelMixin(torch.nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        for p in self.parameters():
            return p.dtype
        return torch.float32

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    def enable_gradient_checkpointing(self):
        self.apply(lambda m: self._set_gradient_checkpointing(m, True))
