"""Layer / Variable / symbolic tensors / functional Model / Sequential (see the package docstring)."""
import collections
import itertools

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ eager tensors
class _Shape(list):
    def as_list(self):
        return list(self)


class ET(torch.Tensor):
    """torch tensor with the one TensorFlow-ism the reference uses (`x.get_shape().as_list()`, diffusion_model.py:72)."""

    def get_shape(self):
        return _Shape(self.shape)


def et(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    elif t.dtype == torch.float64:
        t = t.to(torch.float32)  # Keras casts float inputs to the model's compute dtype
    return t.as_subclass(ET)


# ------------------------------------------------------------------------------------------------ symbolic tensors
def flatten(x):
    if isinstance(x, (list, tuple)):
        for i in x:
            yield from flatten(i)
    elif isinstance(x, dict):
        for k in x:
            yield from flatten(x[k])
    else:
        yield x


def tree_map(fn, x):
    if isinstance(x, (list, tuple)):
        return type(x)(tree_map(fn, i) for i in x)
    if isinstance(x, dict):
        return {k: tree_map(fn, v) for k, v in x.items()}
    return fn(x)


class Node:
    _ids = itertools.count()

    def __init__(self, operation, args, kwargs, is_input=False):
        self.id = next(Node._ids)
        self.operation = operation
        self.args, self.kwargs = args, kwargs
        self.is_input = is_input
        self.input_tensors = [t for t in flatten((args, kwargs)) if isinstance(t, Sym)]
        self.parent_nodes = [t.node for t in self.input_tensors]
        self.outputs = None
        operation._inbound_nodes.append(self)


class Sym:
    """Symbolic tensor of the functional API: carries a dummy value (batch 1) for shape inference while the graph is
    being built, and the node that produces it."""

    def __init__(self, value, node):
        self.value = value
        self.node = node
        self._shape = (None,) + tuple(value.shape[1:])

    @property
    def shape(self):
        return self._shape

    def __add__(self, other):
        return Add()(self, other)

    __radd__ = __add__


class Operation:
    def __init__(self, name=None):
        self.name = name or type(self).__name__.lower()
        self._inbound_nodes = []

    def __call__(self, *args, **kwargs):
        if any(isinstance(t, Sym) for t in flatten((args, kwargs))):
            return self._symbolic_call(args, kwargs)
        return self._eager_call(*args, **kwargs)

    def _symbolic_call(self, args, kwargs):
        unwrap = lambda t: t.value if isinstance(t, Sym) else t  # noqa: E731
        with torch.no_grad():
            out = self._eager_call(*tree_map(unwrap, args), **tree_map(unwrap, kwargs))
        node = Node(self, args, kwargs)
        node.outputs = tree_map(lambda t: Sym(t, node), out)
        return node.outputs

    def _eager_call(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class Add(Operation):  # `x + y` on symbolic tensors (diffusion_model.py:231-234, control_net.py:56)
    def call(self, a, b):
        return a + b


# ------------------------------------------------------------------------------------------------ variables, layers
class Variable:
    def __init__(self, shape, name, dtype=torch.float32):
        self.shape = tuple(int(s) for s in shape)
        self.name = name
        self.value = torch.zeros(self.shape, dtype=dtype)

    def assign(self, w):
        w = torch.as_tensor(np.asarray(w))
        if tuple(w.shape) != self.shape:
            raise ValueError(f"variable {self.name}: shape {self.shape} cannot take a value of shape {tuple(w.shape)}")
        self.value = w.to(self.value.dtype).contiguous()

    def numpy(self):
        return self.value.numpy()


class Layer(Operation):
    _auto = collections.Counter()

    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        if name is None:
            base = type(self).__name__.lower()
            Layer._auto[base] += 1
            name = f"{base}_{Layer._auto[base]}"
        object.__setattr__(self, "_layers", [])
        object.__setattr__(self, "_variables", [])
        Operation.__init__(self, name)
        self.built = False
        self.compute_dtype = "float32"

    def __setattr__(self, key, value):
        # attribute tracking: sub-layers (also inside lists / tuples) are registered in assignment order
        for v in flatten(value) if isinstance(value, (list, tuple, dict)) else (value,):
            if isinstance(v, Layer) and all(v is not l for l in self._layers):
                self._layers.append(v)
        object.__setattr__(self, key, value)

    def add_weight(self, shape, name=None, dtype=torch.float32, **kwargs):
        v = Variable(shape, f"{self.name}/{name}", dtype)
        self._variables.append(v)
        return v

    @property
    def weights(self):
        out = list(self._variables)
        for l in self._layers:
            for v in l.weights:
                if all(v is not o for o in out):
                    out.append(v)
        return out

    def build(self, input_shape):
        pass

    def _eager_call(self, *args, **kwargs):
        if not self.built:
            first = args[0] if args else None
            self.build(tree_map(lambda t: (None,) + tuple(t.shape[1:]) if isinstance(t, torch.Tensor) else t, first))
            self.built = True
        return self.call(*args, **kwargs)

    def call(self, *args, **kwargs):
        raise NotImplementedError


class InputLayer(Layer):
    def call(self):
        raise RuntimeError("InputLayer is a graph source")


def Input(shape=None, batch_size=None, dtype=None, name=None, **kwargs):
    """keras.Input / layers.Input: a symbolic tensor.  Unknown dimensions get a small dummy size for shape inference."""
    layer = InputLayer(name=name)
    dims = [16 if d is None else int(d) for d in shape]
    if len(shape) == 2 and shape[0] is None:
        dims[0] = 77
    is_int = dtype is not None and "int" in str(dtype)
    value = et(torch.zeros([1] + dims, dtype=torch.int64 if is_int else torch.float32))
    node = Node(layer, (), {}, is_input=True)
    sym = Sym(value, node)
    sym._shape = (None,) + tuple(shape)
    sym.dtype = "int32" if is_int else "float32"
    node.outputs = sym
    return sym


# ------------------------------------------------------------------------------------------------ graph ordering
def map_graph(inputs, outputs):
    """Operations of the graph in Keras' order (keras/src/ops/function.py: `map_graph`, `_build_map`): depth is counted
    from the outputs, operations are sorted by decreasing depth, ties by depth-first traversal order from the outputs."""
    finished, in_progress, nodes_in_decreasing_depth, op_index = set(), set(), [], {}
    flat_inputs = list(flatten(inputs))

    def visit(tensor):
        node = tensor.node
        if node in finished:
            return
        if node in in_progress:
            raise ValueError("cycle in the graph")
        op_index.setdefault(node.operation, len(op_index))
        in_progress.add(node)
        if not node.is_input and all(tensor is not t for t in flat_inputs):
            for t in node.input_tensors:
                visit(t)
        finished.add(node)
        in_progress.discard(node)
        nodes_in_decreasing_depth.append(node)

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 20000))
    try:
        for out in flatten(outputs):
            visit(out)
    finally:
        sys.setrecursionlimit(old)
    node_depth, op_depth = {}, {}
    for node in reversed(nodes_in_decreasing_depth):
        depth = node_depth.setdefault(node, 0)
        depth = max(depth, op_depth.get(node.operation, 0))
        op_depth[node.operation] = depth
        node_depth[node] = depth
        for dep in node.parent_nodes:
            node_depth[dep] = max(depth + 1, node_depth.get(dep, 0))
    for t in flat_inputs:
        if t.node.operation not in op_depth:  # an input that no output depends on
            op_depth[t.node.operation] = 0
            op_index[t.node.operation] = -1
    by_depth = collections.defaultdict(list)
    for op, depth in op_depth.items():
        by_depth[depth].append(op)
    ordered = []
    for depth in sorted(by_depth, reverse=True):
        ordered.extend(sorted(by_depth[depth], key=lambda o: op_index[o]))
    nodes = sorted(nodes_in_decreasing_depth, key=lambda n: n.id)  # creation order is a valid execution order
    return ordered, nodes


class Model(Layer):
    """Functional model: `Model(inputs, outputs, name=None)`."""

    def __init__(self, inputs=None, outputs=None, name=None, **kwargs):
        super().__init__(name=name)
        if inputs is None or outputs is None:
            raise NotImplementedError("the shim only supports functional models")
        self._init_graph(inputs, outputs)

    def _init_graph(self, inputs, outputs):
        object.__setattr__(self, "inputs", list(flatten(inputs)))
        object.__setattr__(self, "outputs", outputs)
        ops_, nodes = map_graph(inputs, outputs)
        object.__setattr__(self, "_operations", ops_)
        object.__setattr__(self, "_nodes", nodes)
        object.__setattr__(self, "_layers", [o for o in ops_ if isinstance(o, Layer)])
        self.built = True
        for n in nodes:  # the dummy activations were only needed while the graph was traced
            for t in flatten(n.outputs):
                t.value = None

    @property
    def layers(self):
        return list(self._layers)

    def set_weights(self, weights):
        mine = self.weights
        if len(mine) != len(weights):
            raise ValueError(f"{self.name}: set_weights got {len(weights)} arrays for {len(mine)} variables")
        for v, w in zip(mine, weights):
            v.assign(w)

    def get_weights(self):
        return [v.numpy() for v in self.weights]

    def compile(self, *args, **kwargs):
        pass

    def _run(self, xs):
        if not isinstance(xs, (list, tuple)):
            xs = [xs]
        if len(xs) != len(self.inputs):
            raise ValueError(f"{self.name}: expected {len(self.inputs)} inputs, got {len(xs)}")
        vals = {}
        for sym, x in zip(self.inputs, xs):
            vals[id(sym)] = et(x, torch.int64 if getattr(sym, "dtype", "") == "int32" else torch.float32)
        fetch = lambda t: vals[id(t)] if isinstance(t, Sym) else t  # noqa: E731
        needed = {id(t) for t in flatten(self.outputs)}
        last_use = {}
        for n in self._nodes:
            for t in n.input_tensors:
                last_use[id(t)] = n.id
        with torch.no_grad():
            for n in self._nodes:
                if n.is_input:
                    continue
                out = n.operation._eager_call(*tree_map(fetch, n.args), **tree_map(fetch, n.kwargs))
                for sym, v in zip(flatten(n.outputs), flatten(out)):
                    vals[id(sym)] = v
                for t in n.input_tensors:  # free activations nobody will read again
                    if last_use.get(id(t)) == n.id and id(t) not in needed:
                        vals.pop(id(t), None)
        return tree_map(lambda t: vals[id(t)], self.outputs)

    def predict_on_batch(self, x):
        out = self._run(x)
        return tree_map(lambda t: t.detach().to(torch.float32).numpy() if t.is_floating_point() else t.numpy(), out)

    def call(self, x):
        return self._run(x)


class Sequential(Model):
    """`Sequential([Input(...), layer, ...], name=None)`: layers (and weights) in list order."""

    def __init__(self, layers=None, name=None, **kwargs):
        Layer.__init__(self, name=name)
        layers = list(layers or [])
        if not layers or not isinstance(layers[0], Sym):
            raise NotImplementedError("the shim needs the Input tensor as the first list entry, as the reference passes it")
        x = inp = layers[0]
        for l in layers[1:]:
            x = l(x)
        self._init_graph(inp, x)
        object.__setattr__(self, "_layers", [l for l in layers[1:]])
