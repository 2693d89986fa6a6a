"""Drop-in for models/EODM.py of eastonYi/Unsupervised-ASR (TF2): same names and
signatures (`P_Ngram(kernel, args)`, `EODM_loss(_logits, mask, conv_op, k, py)`),
computed by the eodm_b200 custom ops.  main_EODM.py is used unchanged.

Needs libeodm_tf.so built from tf_shim/eodm_tf_ops.cc against the installed
TensorFlow.  In this repository's image TensorFlow cannot be
installed: the C++ side is compile-checked against stand-in headers
(`make -C tf_shim check`), this module is not executed.  GPU only -- there is no
CPU kernel.  Imports TensorFlow and numpy only (no PyTorch).
"""
import itertools
import os

import numpy as np
import tensorflow as tf

_ops = tf.load_op_library(os.environ.get("EODM_TF_LIB", os.path.join(os.path.dirname(__file__), "..", "libeodm_tf.so")))
_table_ids = itertools.count(1)     # one id per P_Ngram instance: the key of the compact device table in the op library


class _PNgram:
    """Callable like the Keras Model the reference returns; `.summary()` as used at main_EODM.py:63."""

    name = "P_ngram"

    def __init__(self, kernel, args):
        self.kernel = tf.constant(np.asarray(kernel, dtype=np.float32))   # frozen, as trainable=False in the reference
        self.args = args
        self.table_id = next(_table_ids)

    def __call__(self, px):
        @tf.custom_gradient
        def f(x):
            p = _ops.eodm_ngram_prob(x, self.kernel, table_id=self.table_id)
            return p, lambda dp: _ops.eodm_ngram_prob_grad(x, self.kernel, dp, table_id=self.table_id)
        return f(px)

    def counts(self, px, mask):
        @tf.custom_gradient
        def f(x):
            s, n = _ops.eodm_counts(x, mask, self.kernel, table_id=self.table_id)
            return (s, n), lambda gs, gn: _ops.eodm_counts_grad(x, mask, self.kernel, gs, table_id=self.table_id)
        return f(px)

    def loss(self, logits, mask, py):
        """EODM_loss and its gradient as ONE op (softmax, counts, loss, both VJPs): the step bench.py times."""
        @tf.custom_gradient
        def f(x):
            loss, dlogits = _ops.eodm_loss(x, mask, self.kernel, py, table_id=self.table_id)
            return loss, lambda g: g * dlogits
        return f(logits)

    def summary(self):
        n, V, K = self.kernel.shape
        print('Model: "P_ngram"\nconv1d (Conv1D)   (None, None, %d)   %d\nTotal params: %d\nTrainable params: 0\n'
              'Non-trainable params: %d' % (K, n * V * K, n * V * K, n * V * K))


def P_Ngram(kernel, args):
    return _PNgram(kernel, args)


def EODM_loss(_logits, mask, conv_op, k, py):
    """models/EODM.py:5-25 of the reference, without materialising [B, T', K] or the tiled mask."""
    if isinstance(conv_op, _PNgram):
        return conv_op.loss(_logits, tf.cast(mask, tf.bool), tf.convert_to_tensor(py, tf.float32))
    # any other callable: the reference's literal expression
    px_batch = tf.nn.softmax(_logits)
    m = tf.tile(tf.cast(mask, dtype=tf.float32)[:, :, None], [1, 1, k])
    pz = conv_op(px_batch)
    pz = tf.reduce_sum(tf.reduce_sum(pz * m[:, :pz.shape[1], :], 0), 0) / tf.reduce_sum(tf.reduce_sum(m, 0), 0)
    return tf.reduce_sum(-py * tf.math.log(pz + 1e-15))
