"""A STRUCTURAL stand-in for the parts of Keras 2.0.4 that `/root/reference/model.py` touches, so that the
reference's own `omni_model.__init__` and weight-transfer helpers (model.py:33-170) can be executed in the
build container, where Keras / TensorFlow are not installable. Test infrastructure for
`make_transfer_golden.py`; never imported by the product.

It models the graph only: symbolic tensors carry shapes, layers get Keras' auto-names (`dense_1`, ...), a
`Model` lists its layers in topological order, `Dense` holds a [fan_in, units] kernel and a [units] bias that
`get_weights` / `set_weights` / `trainable` act on. No arithmetic is implemented."""
import collections
import sys
import types

import numpy as np

_counters = collections.Counter()


def reset_names():
    _counters.clear()


class Tensor(object):
    def __init__(self, shape, layer, parents=()):
        self.shape, self.layer, self.parents = tuple(shape), layer, tuple(parents)


class Layer(object):
    prefix = "layer"

    def __init__(self, **kw):
        _counters[self.prefix] += 1
        self.name = kw.get("name") or "%s_%d" % (self.prefix, _counters[self.prefix])
        self.trainable = True
        self.input_shape = self.output_shape = None
        self.kwargs = kw

    def get_config(self):
        return {"name": self.name}

    def get_weights(self):
        return []

    def set_weights(self, weights):
        if len(weights) != len(self.get_weights()):
            raise ValueError("layer %s expects %d weights, got %d" % (self.name, len(self.get_weights()), len(weights)))

    def compute_shape(self, shapes):
        return shapes

    def __call__(self, x):
        many = isinstance(x, (list, tuple))
        self.input_shape = [t.shape for t in x] if many else x.shape
        self.output_shape = self.compute_shape(self.input_shape)
        return Tensor(self.output_shape, self, x if many else (x,))


class InputLayer(Layer):
    prefix = "input"


def Input(shape=None, sparse=False, **kw):
    layer = InputLayer(sparse=sparse, **kw)
    layer.input_shape = layer.output_shape = (None,) + tuple(shape)
    return Tensor(layer.output_shape, layer)


class Dense(Layer):
    prefix = "dense"

    def __init__(self, units, activation=None, W_regularizer=None, kernel_regularizer=None, **kw):
        super(Dense, self).__init__(**kw)
        self.units, self.activation = units, activation
        self.kernel_regularizer = kernel_regularizer if kernel_regularizer is not None else W_regularizer
        self.kernel = self.bias = None

    def compute_shape(self, shape):
        self.kernel = np.zeros((shape[-1], self.units), dtype=np.float32)
        self.bias = np.zeros((self.units,), dtype=np.float32)
        return (shape[0], self.units)

    def get_weights(self):
        return [self.kernel, self.bias]

    def set_weights(self, weights):
        k, b = weights
        if k.shape != self.kernel.shape or b.shape != self.bias.shape:
            raise ValueError("Layer weight shape %s not compatible with provided weight shape %s" % (self.kernel.shape, k.shape))
        self.kernel, self.bias = np.array(k), np.array(b)


class Dropout(Layer):
    prefix = "dropout"

    def __init__(self, rate, noise_shape=None, **kw):
        super(Dropout, self).__init__(**kw)
        self.rate, self.noise_shape = rate, noise_shape


class Concatenate(Layer):
    prefix = "concatenate"

    def compute_shape(self, shapes):
        return (shapes[0][0], sum(s[-1] for s in shapes))


class Multiply(Layer):
    prefix = "multiply"

    def compute_shape(self, shapes):
        return shapes[0]


class Lambda(Layer):
    prefix = "lambda"


def concatenate(tensors, **kw):
    return Concatenate(**kw)(tensors)


def multiply(tensors, **kw):
    return Multiply(**kw)(tensors)


class Model(object):
    def __init__(self, inputs=None, outputs=None):
        self.inputs, self.outputs = list(inputs), list(outputs)
        # Keras lists layers by decreasing depth from the outputs; for this chain-shaped graph that is a
        # topological order with the Input layers first, in creation order
        order, seen = [], set()

        def visit(t):
            for p in t.parents:
                visit(p)
            if id(t.layer) not in seen:
                seen.add(id(t.layer))
                order.append(t.layer)

        for t in self.outputs:
            visit(t)
        ins = [l for l in order if isinstance(l, InputLayer)]
        ins.sort(key=lambda l: int(l.name.rsplit("_", 1)[1]))
        self.layers = ins + [l for l in order if not isinstance(l, InputLayer)]

    def get_weights(self):
        return [w for l in self.layers for w in l.get_weights()]

    def set_weights(self, weights):
        weights = list(weights)
        for l in self.layers:
            n = len(l.get_weights())
            l.set_weights(weights[:n])
            weights = weights[n:]


class _Regularizer(object):
    def __init__(self, kind, value):
        self.kind, self.value = kind, value


def install():
    """Put the stub modules into sys.modules under the names model.py imports."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Interfaces(object):
        @staticmethod
        def legacy_dense_support(fn):
            return fn

    keras = mod("keras")
    keras.layers = mod("keras.layers", Input=Input, Dense=Dense, multiply=multiply, Lambda=Lambda,
                       concatenate=concatenate, Dropout=Dropout)
    keras.models = mod("keras.models", Model=Model)
    keras.regularizers = mod("keras.regularizers", l1=lambda v: _Regularizer("l1", v), l2=lambda v: _Regularizer("l2", v))
    keras.legacy = mod("keras.legacy", interfaces=_Interfaces)
    keras.engine = mod("keras.engine", Layer=Layer, InputSpec=object)
    for name in ("activations", "initializers", "constraints", "backend"):
        setattr(keras, name, mod("keras." + name))
    if "tensorflow" not in sys.modules:
        mod("tensorflow", SparseTensor=object)
