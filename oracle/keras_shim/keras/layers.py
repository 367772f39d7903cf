"""The Keras layers the reference instantiates, restated on torch (NHWC in, NHWC out)."""
import torch
import torch.nn.functional as F

from . import activations as _act
from .core import Input, InputLayer, Layer, et  # noqa: F401


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Dense(Layer):
    """y = activation(x @ kernel + bias), kernel (in, units)."""

    def __init__(self, units, activation=None, use_bias=True, **kwargs):
        super().__init__(**kwargs)
        self.units, self.use_bias = int(units), use_bias
        self.activation = _act.get(activation)

    def build(self, input_shape):
        self.kernel = self.add_weight((input_shape[-1], self.units), "kernel")
        self.bias = self.add_weight((self.units,), "bias") if self.use_bias else None

    def call(self, x):
        y = torch.matmul(x, self.kernel.value)
        if self.bias is not None:
            y = y + self.bias.value
        return self.activation(y) if self.activation is not None else y


class Conv2D(Layer):
    """channels_last, padding='valid', kernel (kh, kw, in, filters)."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, **kwargs):
        super().__init__(**kwargs)
        if padding != "valid":
            raise NotImplementedError("the reference only uses padding='valid' (explicit ZeroPadding2D in front)")
        self.filters, self.kernel_size, self.strides, self.use_bias = int(filters), _pair(kernel_size), _pair(strides), use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight(self.kernel_size + (input_shape[-1], self.filters), "kernel")
        self.bias = self.add_weight((self.filters,), "bias") if self.use_bias else None

    def call(self, x):
        w = self.kernel.value.permute(3, 2, 0, 1)  # HWIO -> OIHW
        y = F.conv2d(x.permute(0, 3, 1, 2), w, self.bias.value if self.bias is not None else None, stride=self.strides)
        return y.permute(0, 2, 3, 1)


class ZeroPadding2D(Layer):
    """padding: int | (sym_h, sym_w) | ((top, bottom), (left, right))."""

    def __init__(self, padding=(1, 1), **kwargs):
        super().__init__(**kwargs)
        if isinstance(padding, int):
            padding = ((padding, padding), (padding, padding))
        else:
            h, w = padding
            padding = (_pair(h), _pair(w))
        self.padding = padding

    def call(self, x):
        (t, b), (l, r) = self.padding
        return F.pad(x, (0, 0, l, r, t, b))


class GroupNormalization(Layer):
    """groups=32, axis=-1: moments over (H, W, channels of the group), biased variance, (x - mean) * rsqrt(var + eps)."""

    def __init__(self, groups=32, axis=-1, epsilon=1e-3, center=True, scale=True, **kwargs):
        super().__init__(**kwargs)
        self.groups, self.epsilon = groups, epsilon

    def build(self, input_shape):
        self.gamma = self.add_weight((input_shape[-1],), "gamma")
        self.beta = self.add_weight((input_shape[-1],), "beta")

    def call(self, x):
        shape = x.shape
        C, G = shape[-1], self.groups
        g = x.reshape(shape[0], -1, G, C // G)
        mean = g.mean(dim=(1, 3), keepdim=True)
        var = ((g - mean) ** 2).mean(dim=(1, 3), keepdim=True)
        y = ((g - mean) * torch.rsqrt(var + self.epsilon)).reshape(shape)
        return y * self.gamma.value + self.beta.value


class LayerNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        self.gamma = self.add_weight((input_shape[-1],), "gamma")
        self.beta = self.add_weight((input_shape[-1],), "beta")

    def call(self, x):
        mean = x.mean(dim=-1, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
        return (x - mean) * torch.rsqrt(var + self.epsilon) * self.gamma.value + self.beta.value


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.fn = _act.get(activation)

    def call(self, x):
        return self.fn(x)


class UpSampling2D(Layer):
    def __init__(self, size=(2, 2), interpolation="nearest", **kwargs):
        super().__init__(**kwargs)
        self.size = _pair(size)

    def call(self, x):
        return x.repeat_interleave(self.size[0], dim=1).repeat_interleave(self.size[1], dim=2)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, xs):
        return torch.cat(list(xs), dim=self.axis)


class Rescaling(Layer):
    def __init__(self, scale, offset=0.0, **kwargs):
        super().__init__(**kwargs)
        self.scale, self.offset = scale, offset

    def call(self, x):
        return x * self.scale + self.offset


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super().__init__(**kwargs)
        self.function = function

    def call(self, x):
        return self.function(x)


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, **kwargs):
        super().__init__(**kwargs)
        self.input_dim, self.output_dim = input_dim, output_dim

    def build(self, input_shape):
        self.embeddings = self.add_weight((self.input_dim, self.output_dim), "embeddings")

    def call(self, ids):
        return et(self.embeddings.value[ids.long()])
