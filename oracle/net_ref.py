"""numpy restatement, layer by layer, of the reference's policy/value network (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/custom_alphazero/model/tensorflow/base_layers.py:20-125 (InnerConvBlock, OuterConvBlock) and
model/tensorflow/model.py:21-188 (ResidualTower, PolicyHead, ValueHead, PolicyValueModel.call) in Keras' own
conventions: NHWC activations, HWIO convolution kernels, [in, out] dense kernels, `Flatten` over (h, w, c),
BatchNormalization in inference mode with Keras' default epsilon 1e-3.  Nothing here shares code or layout with
az_b200/net.py (which is PyTorch, NCHW-logical, OIHW kernels, [out, in] dense weights, BN folded for inference).

The weights come in as the list `PolicyValueModel.get_weights()` returns.  Its order is restated from Keras 2.7.0
(poetry.lock pins tensorflow 2.7.1 / keras 2.7.0): `Model.weights` concatenates `layer.weights` of the model's tracked
layers in attribute order (residual_tower, policy_head, value_head - model.py:156-160), and a nested `Layer.weights`
is all its trainable variables in depth-first layer order followed by all its non-trainable ones (the BatchNormalization
moving statistics).  Layers are tracked in the order their attributes are first bound to a Layer:
  InnerConvBlock (base_layers.py:33-57): conv_layer (kernel, bias), batch_normalization_layer (gamma, beta | mean, variance)
  OuterConvBlock (:85-115):             inner_conv_1, inner_conv_2, residual_connexion
  ResidualTower (model.py:34-58):       conv_blocks[0] = InnerConvBlock stem, conv_blocks[1..depth] = OuterConvBlock
  PolicyHead (:82-96):                  inner_conv, dense            ValueHead (:124-143): inner_conv, dense_1, dense_2

Parity status of THIS file: UNPINNED against TensorFlow (not installed here, no checkpoint ships with the reference).
What it pins is the architecture of az_b200.net.PolicyValueNet by an independent restatement of the reference's layer
definitions: tests/test_net_keras_restatement.py feeds both the same weights and positions."""
import numpy as np

BN_EPSILON = 1e-3  # tf.keras.layers.BatchNormalization default (base_layers.py:55-57 passes no epsilon)


def conv2d_same(x, kernel, bias):
    """Conv2D(padding="same", strides=(1, 1)) (base_layers.py:35-51): x [n, h, w, cin], kernel [kh, kw, cin, cout]
    (cross-correlation, as TensorFlow computes it), bias [cout]."""
    n, h, w, cin = x.shape
    kh, kw, kcin, cout = kernel.shape
    assert kcin == cin and kh % 2 == 1 and kw % 2 == 1
    ph, pw = kh // 2, kw // 2
    xp = np.zeros((n, h + 2 * ph, w + 2 * pw, cin), dtype=np.float64)
    xp[:, ph:ph + h, pw:pw + w, :] = x
    out = np.zeros((n, h, w, cout), dtype=np.float64)
    for dy in range(kh):
        for dx in range(kw):
            out += xp[:, dy:dy + h, dx:dx + w, :] @ kernel[dy, dx].astype(np.float64)
    return out + bias.astype(np.float64)


def batch_normalization(x, gamma, beta, moving_mean, moving_variance):
    """BatchNormalization, training=False: gamma * (x - mean) / sqrt(variance + epsilon) + beta over the channel axis."""
    return (x - moving_mean) / np.sqrt(moving_variance.astype(np.float64) + BN_EPSILON) * gamma + beta


def relu(x):
    return np.maximum(x, 0.0)


class _Take:
    """Hands out the arrays of one `Layer.weights` list: trainable variables first, then the non-trainable ones."""

    def __init__(self, weights, n_trainable):
        self.t = list(weights[:n_trainable])
        self.nt = list(weights[n_trainable:])

    def conv_bn(self):
        kernel, bias, gamma, beta = (self.t.pop(0) for _ in range(4))
        mean, variance = self.nt.pop(0), self.nt.pop(0)
        return kernel, bias, gamma, beta, mean, variance

    def dense(self):
        return self.t.pop(0), self.t.pop(0)


def inner_conv_block(x, p, activation):
    """base_layers.py:59-67: conv, then batch normalisation, then the activation (if any)."""
    kernel, bias, gamma, beta, mean, variance = p
    y = batch_normalization(conv2d_same(x, kernel, bias), gamma, beta, mean, variance)
    return relu(y) if activation else y


def outer_conv_block(x, p1, p2, pr):
    """base_layers.py:117-125: inner_conv_1 (relu) -> inner_conv_2 (no activation); the shortcut is a 1x1 InnerConvBlock
    with batch normalisation on the block's INPUT (a projection, not an identity: :101-112); Add; relu."""
    y = inner_conv_block(inner_conv_block(x, p1, True), p2, False)
    return relu(inner_conv_block(x, pr, False) + y)


def split_weights(weights, depth):
    """get_weights() list -> (tower list, policy-head list, value-head list) with their trainable counts."""
    n_tower_t, n_tower = 4 + 12 * depth, 6 + 18 * depth
    tower = _Take(weights[:n_tower], n_tower_t)
    policy = _Take(weights[n_tower:n_tower + 8], 6)
    value = _Take(weights[n_tower + 8:n_tower + 18], 8)
    assert len(weights) == n_tower + 18, (len(weights), n_tower + 18)
    return tower, policy, value


def residual_tower(x, tower, depth):
    """model.py:60-65: the stem InnerConvBlock (3x3, relu, BN) and `depth` OuterConvBlocks."""
    stem = tower.conv_bn()
    blocks = [(tower.conv_bn(), tower.conv_bn(), tower.conv_bn()) for _ in range(depth)]
    y = inner_conv_block(x, stem, True)
    for p1, p2, pr in blocks:
        y = outer_conv_block(y, p1, p2, pr)
    return y


def policy_head(x, head):
    """model.py:98-104: 1x1 InnerConvBlock with 2 filters (relu, BN), Flatten over (h, w, c), Dense(A, softmax)."""
    conv = head.conv_bn()
    kernel, bias = head.dense()
    f = inner_conv_block(x, conv, True).reshape(x.shape[0], -1)
    logits = f @ kernel.astype(np.float64) + bias
    e = np.exp(logits - logits.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def value_head(x, head):
    """model.py:145-152: 1x1 InnerConvBlock with 1 filter (relu, BN), Flatten, Dense(256, relu), Dense(1, tanh)."""
    conv = head.conv_bn()
    k1, b1 = head.dense()
    k2, b2 = head.dense()
    f = inner_conv_block(x, conv, True).reshape(x.shape[0], -1)
    return np.tanh(relu(f @ k1.astype(np.float64) + b1) @ k2.astype(np.float64) + b2)


def policy_value_model(states, weights, depth=4, return_tower=False):
    """PolicyValueModel.call (model.py:182-188) on `states` [n, h, w, planes] with the get_weights() list `weights`:
    -> (policy [n, A], value [n, 1]) in float64 (Keras computes in float32; the callers compare with a tolerance)."""
    tower, policy, value = split_weights(list(weights), depth)
    t = residual_tower(np.asarray(states, dtype=np.float64), tower, depth)
    out = policy_head(t, policy), value_head(t, value)
    return out + (t,) if return_tower else out


def n_parameters(weights, depth=4):
    """Trainable parameter count (what model.summary() calls 'Trainable params')."""
    tower, policy, value = split_weights(list(weights), depth)
    return sum(int(np.prod(a.shape)) for part in (tower, policy, value) for a in part.t)
