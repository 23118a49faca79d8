"""The stand-in `tensorflow` module (oracle/tf_shim) underpins every tests/golden/ref_* fixture: the
reference's own Python ran on it.  These tests pin the tf semantics the hot path relies on, each
against the behaviour documented for the TensorFlow op of the same name (cited per case), so that a
difference between the stand-in and TensorFlow would have to be a difference in IEEE arithmetic, not in
op semantics.  Runs in a subprocess: importing the stand-in patches a few torch.Tensor methods."""

import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = textwrap.dedent('''
    import math, sys
    import numpy as np
    sys.path.insert(0, sys.argv[1])
    import tensorflow as tf
    assert tf.__version__ == "2.shim"
    T = lambda a: tf.constant(a, dtype=tf.float64)

    # tf.math.mod / floormod: result has the sign of the divisor (tf.math.floormod docs)
    assert np.allclose(tf.math.mod(T([-0.5, 7.5, -7.5]), 2 * math.pi).numpy(),
                       np.mod([-0.5, 7.5, -7.5], 2 * math.pi))
    # tf.math.l2_normalize: x * rsqrt(max(sum(x^2), epsilon)), epsilon = 1e-12
    x = T([[3.0, 4.0, 0.0], [0.0, 0.0, 0.0], [1e-9, 0.0, 0.0]])
    n = tf.math.l2_normalize(x, axis=1).numpy()
    assert np.allclose(n[0], [0.6, 0.8, 0.0]) and np.all(n[1] == 0.0) and np.allclose(n[2, 0], 1e-9 / 1e-6)
    # tf.linalg.normalize returns (x / ||x||, ||x||) along the axis
    v, nrm = tf.linalg.normalize(T([[3.0, 4.0, 0.0]]), axis=1)
    assert np.allclose(v.numpy(), [[0.6, 0.8, 0.0]]) and np.allclose(nrm.numpy(), [[5.0]])
    # tf.meshgrid default indexing "xy": outputs have shape (len(second), len(first))
    a, b = tf.meshgrid(T([1.0, 2.0, 3.0]), T([10.0, 20.0]))
    assert tuple(a.shape) == (2, 3) and np.array_equal(a.numpy()[1], [1, 2, 3]) and np.array_equal(b.numpy()[:, 0], [10, 20])
    # tf.argmin: first minimal index on ties (Eigen CPU kernel), int64 output
    am = tf.argmin(T([[2.0, 1.0, 5.0], [1.0, 1.0, 5.0], [1.0, 3.0, 5.0]]), axis=0)
    assert am.dtype == tf.int64 and am.numpy().tolist() == [1, 0, 0]
    # tf.where(cond, x, y) selects; scalars broadcast
    assert tf.where(T([1.0, -1.0]) < 0, T([5.0, 6.0]), 0.0).numpy().tolist() == [0.0, 6.0]
    # tf.gather_nd with (N, 2) indices picks elements; tf.gather along axis 0; tf.boolean_mask keeps order
    m = T([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    assert tf.gather_nd(m, tf.constant([[2, 0], [0, 1]], dtype=tf.int64)).numpy().tolist() == [5.0, 2.0]
    assert tf.gather(m, tf.constant([2, 0], dtype=tf.int64)).numpy().tolist() == [[5.0, 6.0], [1.0, 2.0]]
    assert tf.boolean_mask(T([1.0, 2.0, 3.0, 4.0]), tf.constant([True, False, True, True])).numpy().tolist() == [1.0, 3.0, 4.0]
    # tf.reduce_max over everything, tf.reduce_any along an axis
    assert float(tf.reduce_max(m)) == 6.0
    assert tf.reduce_any(tf.constant([[True, False], [False, False]]), axis=0).numpy().tolist() == [True, False]
    # tf.concat / tf.stack / tf.unstack / tf.transpose / tf.pad / tf.repeat
    assert tf.concat([T([1.0]), T([2.0, 3.0])], 0).numpy().tolist() == [1.0, 2.0, 3.0]
    assert tuple(tf.transpose(tf.stack([T([1.0, 2.0]), T([3.0, 4.0])])).shape) == (2, 2)
    assert [u.numpy().tolist() for u in tf.unstack(m, axis=1)] == [[1.0, 3.0, 5.0], [2.0, 4.0, 6.0]]
    assert tf.pad(T([1.0, 2.0]), tf.constant([[1, 2]])).numpy().tolist() == [0.0, 1.0, 2.0, 0.0, 0.0]
    assert tf.repeat(T([1.0, 2.0]), [2, 1]).numpy().tolist() == [1.0, 1.0, 2.0]
    # tf.shape returns a tensor usable as a shape; tf.broadcast_to; tf.rank of a python scalar is 0
    assert tf.broadcast_to(1.5, tf.shape(T([0.0, 0.0, 0.0]))).numpy().tolist() == [1.5, 1.5, 1.5]
    assert tf.rank(3) == 0 and tf.rank(T([1.0])) == 1
    # tf.range dtype follows its arguments; tf.linspace includes both ends
    assert tf.range(3).dtype == tf.int32 and tf.range(2, dtype=tf.int64).dtype == tf.int64
    assert np.allclose(tf.linspace(T(0.0), T(1.0), 5).numpy(), [0, 0.25, 0.5, 0.75, 1.0])
    # python sequence on the left of a tensor operator converts (Tensor.__rsub__ in TF)
    assert ((0.0, 0.0, 4.0) - T([1.0, 1.0, 1.0])).numpy().tolist() == [-1.0, -1.0, 3.0]
    # Variable / assign_add / GradientTape.gradient of a non-scalar = gradient of its sum;
    # tf.stop_gradient blocks; tf.where routes the gradient to the selected branch only
    p = tf.Variable(T([1.0, 2.0, 3.0]), dtype=tf.float64)
    with tf.GradientTape() as tape:
        y = tf.where(p > 1.5, p * p, tf.stop_gradient(p) * 10.0)
        g = tape.gradient(y, p)
    assert g.numpy().tolist() == [0.0, 4.0, 6.0]
    p.assign_add(T([1.0, 1.0, 1.0]))
    assert p.numpy().tolist() == [2.0, 3.0, 4.0]
    # Keras SGD(nesterov=True): v = m v - lr g ; w += m v - lr g   (default learning rate 0.01)
    opt = tf.optimizers.SGD(nesterov=True)
    opt.momentum = 0.5
    w = tf.Variable(T([1.0]), dtype=tf.float64)
    opt.apply_gradients([(T([2.0]), w)])          # v = -0.02 ; w = 1 + 0.5 * -0.02 - 0.02
    assert abs(float(w) - 0.97) < 1e-15
    opt.apply_gradients([(T([2.0]), w)])          # v = -0.03 ; w += -0.015 - 0.02
    assert abs(float(w) - 0.935) < 1e-15
    print("shim ok")
''')


def test_tf_stand_in_semantics():
    shim = os.path.join(ROOT, "oracle", "tf_shim")
    proc = subprocess.run([sys.executable, "-c", CASES, shim], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0 and "shim ok" in proc.stdout, proc.stdout + proc.stderr
