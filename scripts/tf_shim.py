"""A NumPy-backed stand-in for the slice of TensorFlow / Keras / tf-models / ml_collections that the
reference's MODEL code touches, so that `deepconsensus/models/{networks,encoder_stack,attention_layer,
ffn_layer,data_providers}.py` can be EXECUTED unmodified in this container (no TensorFlow here).

TEST INFRASTRUCTURE (used only by scripts/make_model_golden.py).  What this buys: the oracle
(oracle/model.py) is compared against the reference's own Python -- its op order, concat order, scaling,
masking, residual wiring, variable naming -- rather than against my reading of it.  What it does NOT buy:
the primitives below (Dense, EinsumDense, LayerNormalization, Softmax, tf.cast, band_part, and the two
tf-models layers OnDeviceEmbedding / RelativePositionEmbedding) are themselves restated from their
published behaviour (Keras 2.9, tf-models-official 2.9.1), in float32.
"""
import contextlib
import math
import sys
import types

import numpy as np

F32 = np.float32


class T(np.ndarray):
  """ndarray with the couple of Tensor methods the reference calls."""

  def set_shape(self, shape):
    assert tuple(self.shape) == tuple(shape), (self.shape, shape)

  def numpy(self):
    return np.asarray(self)


def _t(x):
  return np.asarray(x).view(T)


class TensorShape:
  def __init__(self, dims):
    self._dims = list(dims)

  def as_list(self):
    return list(self._dims)

  def __getitem__(self, i):
    return self._dims[i]


# --------------------------------------------------------------------------- keras base classes
class Layer:
  def __init__(self, *args, name=None, dtype=None, **kwargs):
    self._built = False
    self.name = name

  def build(self, input_shape):
    self._built = True

  def __call__(self, *args, **kwargs):
    if not self._built:
      first = args[0] if args else next(iter(kwargs.values()))
      self.build(TensorShape(np.shape(first)))
      self._built = True
    return self.call(*args, **kwargs)


class Model(Layer):
  def summary(self):
    pass


def _activation(a):
  if a is None:
    return lambda x: x
  if callable(a):
    return a
  if a == "relu":
    return relu
  raise NotImplementedError(a)


class Dense(Layer):
  """keras.layers.Dense: y = act(x @ kernel + bias), kernel [in, units]."""

  def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None,
               bias_initializer=None, name=None, **kw):
    super().__init__(name=name)
    self.units, self.use_bias, self.activation = units, use_bias, _activation(activation)
    self.kernel = None
    self.bias = None

  def build(self, input_shape):
    self.kernel = np.zeros((input_shape.as_list()[-1], self.units), F32)
    if self.use_bias:
      self.bias = np.zeros((self.units,), F32)
    super().build(input_shape)

  def call(self, x):
    y = np.matmul(np.asarray(x, F32), self.kernel)
    if self.use_bias:
      y = y + self.bias
    return _t(self.activation(y).astype(F32))


class EinsumDense(Layer):
  """keras.layers.experimental.EinsumDense without bias: y = einsum(equation, x, kernel)."""

  def __init__(self, equation, output_shape, kernel_initializer=None, bias_axes=None, name=None, **kw):
    super().__init__(name=name)
    assert bias_axes is None
    self.equation, self.out_shape = equation, tuple(output_shape)
    self.kernel = None

  def build(self, input_shape):
    lhs, out = self.equation.split("->")
    a, k = lhs.split(",")
    sizes = dict(zip(a, input_shape.as_list()))
    # output_shape omits the batch dimension
    for letter, n in zip(out[1:], self.out_shape):
      if n is not None:
        sizes[letter] = n
    self.kernel = np.zeros([sizes[c] for c in k], F32)
    super().build(input_shape)

  def call(self, x):
    return _t(np.einsum(self.equation, np.asarray(x, F32), self.kernel).astype(F32))


class LayerNormalization(Layer):
  """keras LayerNormalization over the last axis: (x - mean) * rsqrt(var + eps) * gamma + beta."""

  def __init__(self, epsilon=1e-3, dtype=None, name=None, **kw):
    super().__init__(name=name)
    self.epsilon = epsilon
    self.gamma = self.beta = None

  def build(self, input_shape):
    n = input_shape.as_list()[-1]
    self.gamma, self.beta = np.ones((n,), F32), np.zeros((n,), F32)
    super().build(input_shape)

  def call(self, x):
    x = np.asarray(x, F32)
    mean = x.mean(-1, keepdims=True, dtype=F32)
    var = ((x - mean) ** 2).mean(-1, keepdims=True, dtype=F32)
    return _t(((x - mean) * (F32(1) / np.sqrt(var + F32(self.epsilon))) * self.gamma + self.beta).astype(F32))


def softmax(x, axis=-1, name=None):
  x = np.asarray(x, F32)
  e = np.exp(x - x.max(axis, keepdims=True))
  return _t((e / e.sum(axis, keepdims=True)).astype(F32))


class Softmax(Layer):
  def call(self, x):
    return softmax(x)


def relu(x):
  return _t(np.maximum(np.asarray(x), 0))


# --------------------------------------------------------------------------- tf-models layers [3P, restated]
class OnDeviceEmbedding(Layer):
  """official.nlp.modeling.layers.OnDeviceEmbedding: gather(embeddings, ids) * scale_factor."""

  def __init__(self, vocab_size, embedding_width, initializer=None, use_one_hot=False, scale_factor=None,
               name=None, **kw):
    super().__init__(name=name)
    self._vocab_size, self._embedding_width, self._scale_factor = vocab_size, embedding_width, scale_factor
    self.embeddings = np.zeros((vocab_size, embedding_width), F32)

  def call(self, inputs):
    ids = np.asarray(inputs).astype(np.int64)
    if ids.min() < 0 or ids.max() >= self._vocab_size:
      raise IndexError("InvalidArgumentError: embedding id out of range")   # TF CPU gather behaviour
    e = self.embeddings[ids]
    if self._scale_factor:
      e = e * F32(self._scale_factor)
    return _t(e.astype(F32))


class RelativePositionEmbedding(Layer):
  """official.nlp.modeling.layers.RelativePositionEmbedding(hidden_size, min_timescale=1, max_timescale=1e4)."""

  def __init__(self, hidden_size, min_timescale=1.0, max_timescale=1.0e4, name=None, **kw):
    super().__init__(name=name)
    self._hidden_size, self._min, self._max = hidden_size, min_timescale, max_timescale

  def call(self, inputs, length=None):
    length = np.shape(inputs)[1] if length is None else length
    position = np.arange(length, dtype=F32)
    num_timescales = self._hidden_size // 2
    log_inc = F32(math.log(float(self._max) / float(self._min)) / (float(num_timescales) - 1))
    inv = (F32(self._min) * np.exp(np.arange(num_timescales, dtype=F32) * -log_inc)).astype(F32)
    scaled = position[:, None] * inv[None, :]
    return _t(np.concatenate([np.sin(scaled), np.cos(scaled)], axis=1).astype(F32))


# --------------------------------------------------------------------------- module assembly
class ConfigDict(dict):
  def __getattr__(self, k):
    try:
      return self[k]
    except KeyError as e:
      raise AttributeError(k) from e

  def __setattr__(self, k, v):
    self[k] = v


def band_part(x, lower, upper):
  x = np.asarray(x)
  n, m = x.shape[-2:]
  i, j = np.arange(n)[:, None], np.arange(m)[None, :]
  keep = ((lower < 0) | (i - j <= lower)) & ((upper < 0) | (j - i <= upper))
  return _t(np.where(keep, x, 0).astype(x.dtype))


def cast(x, dtype):
  x = np.asarray(x)
  if dtype in ("int32", np.int32):
    return _t(np.trunc(x).astype(np.int32))     # float -> int32 truncates toward zero
  return _t(x.astype(F32))


class _Anything:
  """Permissive placeholder for names only used in type annotations (tf.data.Dataset, tf.train.Example ...)."""

  def __getattr__(self, k):
    if k.startswith("__"):
      raise AttributeError(k)
    return _Anything()

  def __call__(self, *a, **k):
    return _Anything()

  def __getitem__(self, k):
    return _Anything()

  def __mro_entries__(self, bases):
    return (object,)


class _NS(_Anything):
  """Namespace with explicit attributes and a permissive fallback for everything else."""

  def __init__(self, **kw):
    self.__dict__.update(kw)


def install():
  """Registers stub modules: tensorflow, tensorflow.compat.v2, ml_collections, official..., pysam, absl."""
  tf = types.ModuleType("tensorflow")

  def _fallback(name):
    if name.startswith("__"):
      raise AttributeError(name)
    return _Anything()
  tf.__getattr__ = _fallback
  tf.float32, tf.int32, tf.int64, tf.string = "float32", "int32", "int64", "string"
  tf.Tensor, tf.TensorShape = np.ndarray, TensorShape
  tf.cast = cast
  tf.squeeze = lambda x, axis=None: _t(np.squeeze(x, axis))
  tf.transpose = lambda x, perm=None: _t(np.transpose(x, perm))
  tf.zeros_like = lambda x: _t(np.zeros_like(np.asarray(x)))
  tf.zeros = lambda shape, dtype=None: _t(np.zeros(shape, F32))
  tf.ones = lambda shape, dtype=None: _t(np.ones(shape, F32))
  tf.reduce_sum = lambda x, axis=None: _t(np.sum(x, axis=axis))
  tf.expand_dims = lambda x, axis: _t(np.expand_dims(x, axis))
  tf.concat = lambda xs, axis: _t(np.concatenate([np.asarray(x) for x in xs], axis=axis))
  tf.einsum = lambda eq, *ops: _t(np.einsum(eq, *[np.asarray(o, F32) for o in ops]).astype(F32))
  tf.where = lambda c, a, b: _t(np.where(c, a, F32(b) if np.isscalar(b) else b).astype(F32))
  tf.not_equal = lambda a, b: np.not_equal(a, b)
  tf.reshape = lambda x, s: _t(np.reshape(x, s))
  tf.clip_by_value = lambda x, clip_value_min, clip_value_max: _t(np.clip(x, clip_value_min, clip_value_max))
  tf.convert_to_tensor = lambda x: _t(np.asarray(x))
  tf.name_scope = lambda name: contextlib.nullcontext()
  tf.function = lambda f=None, **kw: f if f is not None else (lambda g: g)
  tf.Variable = lambda initial_value=None, trainable=True, **kw: np.array(initial_value, dtype=F32)
  tf.zeros_initializer = lambda: (lambda shape, dtype=None: np.zeros(shape, F32))
  tf.random_normal_initializer = lambda mean=0.0, stddev=1.0: None
  tf.nn = _NS(softmax=softmax, relu=relu, dropout=lambda x, rate: x)
  tf.linalg = types.SimpleNamespace(band_part=band_part)
  tf.io = _NS(FixedLenFeature=lambda *a, **k: None, gfile=None)
  tf.math = _NS(log=np.log)
  keras = _NS()
  keras.Model, keras.Input = Model, None
  keras.layers = _NS(Layer=Layer, Dense=Dense, LayerNormalization=LayerNormalization,
                                       Softmax=Softmax, Flatten=None, Reshape=None, Dropout=None,
                                       experimental=_NS(EinsumDense=EinsumDense))
  keras.initializers = _NS(RandomUniform=lambda minval, maxval: None)
  keras.regularizers = _NS(l2=lambda *a: None)
  keras.applications = _NS()
  tf.keras = keras
  compat = types.ModuleType("tensorflow.compat")
  v2 = types.ModuleType("tensorflow.compat.v2")
  v2.__dict__.update(tf.__dict__)
  compat.v2 = v2
  tf.compat = compat
  sys.modules.update({"tensorflow": tf, "tensorflow.compat": compat, "tensorflow.compat.v2": v2})

  mlc = types.ModuleType("ml_collections")
  mlc.ConfigDict = ConfigDict
  mlc.FrozenConfigDict = ConfigDict
  cd_pkg = types.ModuleType("ml_collections.config_dict")
  cd_mod = types.ModuleType("ml_collections.config_dict.config_dict")
  cd_mod.ConfigDict = cd_mod.FrozenConfigDict = ConfigDict
  cd_pkg.config_dict = cd_mod
  cd_pkg.ConfigDict = cd_pkg.FrozenConfigDict = ConfigDict
  mlc.config_dict = cd_pkg
  sys.modules.update({"ml_collections": mlc, "ml_collections.config_dict": cd_pkg,
                      "ml_collections.config_dict.config_dict": cd_mod})

  layers = types.ModuleType("official.nlp.modeling.layers")
  layers.OnDeviceEmbedding, layers.RelativePositionEmbedding = OnDeviceEmbedding, RelativePositionEmbedding
  for name in ("official", "official.nlp", "official.nlp.modeling"):
    sys.modules[name] = types.ModuleType(name)
  sys.modules["official.nlp.modeling"].layers = layers
  sys.modules["official.nlp.modeling.layers"] = layers

  pysam = types.ModuleType("pysam")
  for i, n in enumerate(["CMATCH", "CINS", "CDEL", "CREF_SKIP", "CSOFT_CLIP", "CHARD_CLIP", "CPAD", "CEQUAL",
                         "CDIFF", "CBACK"]):
    setattr(pysam, n, i)
  sys.modules["pysam"] = pysam
  return tf
