"""Global numeric settings (mirror of reference ``src/misc/settings.py:5-34``): float32 everywhere, jitter 1e-5,
device = the current CUDA device when one is visible. Unlike the reference nothing here relies on
``torch.set_default_tensor_type``; every module creates its tensors on ``settings.device`` explicitly."""
import numpy
import torch


class Settings:
    @property
    def torch_int(self):
        return torch.int32

    @property
    def numpy_int(self):
        return numpy.int32

    @property
    def device(self):
        if torch.cuda.is_available():
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    @property
    def torch_float(self):
        return torch.float32

    @property
    def numpy_float(self):
        return numpy.float32

    @property
    def jitter(self):
        return 1e-5


settings = Settings()
