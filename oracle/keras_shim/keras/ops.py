"""keras.ops subset used by the reference."""
import numpy as np
import torch

from .core import et


def shape(x):
    return tuple(int(s) for s in x.shape)


def reshape(x, newshape):
    return x.reshape(tuple(int(s) for s in newshape))


def transpose(x, axes=None):
    return x.permute(*axes) if axes is not None else x.T


def einsum(subscripts, *operands):
    return torch.einsum(subscripts, *operands)


def cast(x, dtype):
    td = {"float32": torch.float32, "int32": torch.int32, "float16": torch.float16}[str(dtype)]
    return et(torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x, td)


def sqrt(x):
    return torch.sqrt(x if isinstance(x, torch.Tensor) else torch.as_tensor(float(x)))


def sigmoid(x):
    return torch.sigmoid(x)


def split(x, indices_or_sections, axis=0):
    return list(torch.tensor_split(x, indices_or_sections, dim=axis))


class nn:  # noqa: N801  (keras.ops.nn namespace)
    @staticmethod
    def softmax(x, axis=-1):
        return torch.softmax(x, dim=axis)
