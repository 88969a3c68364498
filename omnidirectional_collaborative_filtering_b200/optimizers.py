"""Keras-2.0.4-shaped optimizer descriptions (`train.py:12,50-51`, `train_jester.py:61`).
Only hyper-parameters live here; the update rules run fused in `csrc/`."""
from __future__ import annotations


class Optimizer(object):
    kind = "sgd"

    def __init__(self, lr, epsilon=1e-8, decay=0.0, p1=0.0, p2=0.0):
        self.lr, self.epsilon, self.decay, self.p1, self.p2 = float(lr), float(epsilon), float(decay), float(p1), float(p2)

    def get_config(self):
        return dict(kind=self.kind, lr=self.lr, epsilon=self.epsilon, decay=self.decay, p1=self.p1, p2=self.p2)


class SGD(Optimizer):
    kind = "sgd"

    def __init__(self, lr=0.01, decay=0.0):
        super(SGD, self).__init__(lr, 0.0, decay)


class Adagrad(Optimizer):
    kind = "adagrad"

    def __init__(self, lr=0.01, epsilon=1e-8, decay=0.0):
        super(Adagrad, self).__init__(lr, epsilon, decay)


class RMSprop(Optimizer):
    kind = "rmsprop"

    def __init__(self, lr=0.001, rho=0.9, epsilon=1e-8, decay=0.0):
        super(RMSprop, self).__init__(lr, epsilon, decay, p1=rho)
        self.rho = rho


class Adam(Optimizer):
    kind = "adam"

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-8, decay=0.0):
        super(Adam, self).__init__(lr, epsilon, decay, p1=beta_1, p2=beta_2)
        self.beta_1, self.beta_2 = beta_1, beta_2


_BY_NAME = {"sgd": SGD, "adagrad": Adagrad, "rmsprop": RMSprop, "adam": Adam}


def get(spec) -> Optimizer:
    """'adagrad' | 'rmsprop' | 'adam' | 'sgd' (Keras defaults) or an Optimizer instance."""
    if isinstance(spec, Optimizer):
        return spec
    if isinstance(spec, str) and spec.lower() in _BY_NAME:
        return _BY_NAME[spec.lower()]()
    raise ValueError("unknown optimizer %r" % (spec,))
