"""numpy restatement of the reference's training-step arithmetic (TEST INFRASTRUCTURE ONLY).

Parity status of THIS file: UNPINNED - the reference's trainer is TensorFlow/Keras, which is not
installed here, so no golden vectors could be produced from it; the formulas follow the reference's
definitions (model/tensorflow/base_layers.py:12-17 losses, Keras l2 regulariser, Keras SGD update)."""
import numpy as np

EPS = 1e-7


def policy_loss(pi, p):
    return float(np.mean(np.sum(-pi * np.log(p + EPS), axis=-1)))


def value_loss(z, v):
    return float(np.mean((v - z) ** 2))


def l2_penalty(kernels, l2=1e-4):
    return float(l2 * sum(float((w.astype(np.float64) ** 2).sum()) for w in kernels))


def keras_sgd(w, g, v, lr, momentum=0.9):
    v_new = momentum * v - lr * g
    return w + v_new, v_new
