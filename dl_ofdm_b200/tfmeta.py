"""`.meta` (MetaGraphDef) writer: the graph `dev/py/ofdmreceiver_np.py:121-192` builds, emitted without TensorFlow.

`tf.train.Saver.save` writes three files per checkpoint; the reference restores through the third one
(`tf.train.import_meta_graph(path + '.meta')`, dev/py/model.py:51-56, dev/py/ofdmreceiver_np_mp.py:265,375) and then
fetches tensors BY NAME (`bits_in:0`, `tx_ofdm:0`, `input:0`, `output:0`, `cost:0`, `log_ber:0`, `linear_ber:0`,
`conf_matrix:0`, `tx_power:0`, `noise_power:0`, `iq_rx:0`, `iq_tx:0`, `ce_mean:0`, `SNR:0`).  `write_meta` emits that
graph -- placeholders, the `transmitter` / `channel` / `receiver` scopes op by op as TF's Python library lays them out,
the loss / BER / confusion-matrix tail, the variables with their initializers, and the `save/` sub-graph of a V2
`tf.train.Saver` -- so that a bundle written by `tfbundle.write_checkpoint` can be picked up by the reference scripts.

What is checked and what is not: tests/test_host.py::test_meta_graph_matches_shipped_v1 rebuilds the graph of each of
the eight checkpoints the reference ships (`test_v1/model/*.meta`, TF 1.10.1) and requires every node reachable from
the named fetches, every variable / Assign / read node and the whole `save/` sub-graph to be identical in name, op,
inputs and attributes, and the `variables` / `trainable_variables` collections and `saver_def` to match.  The
optimizer part of the training graph (gradients, Adam slots, `train_op`) is NOT emitted: a restored graph serves
inference and the BER sweeps (`--test=True`), not a continuation of training.  `Saver.restore` itself cannot be run
here (no TensorFlow on this image).  Graphs of the dev architecture (7-symbol LTE frame) follow dev/py/model.py:1222-1292
with the same emitter; no dev `.meta` is shipped to compare them with.
"""
from __future__ import annotations

import numpy as np
from tensorboard.compat.proto import meta_graph_pb2, saver_pb2, variable_pb2

from . import tfgraph as tg
from .tfgraph import FLOAT, INT32, INT64, BOOL, STRING, HALF, T

FETCHES = ('bits_in', 'tx_ofdm', 'SNR', 'input', 'output', 'cost', 'log_ber', 'linear_ber', 'conf_matrix', 'tx_power',
           'noise_power', 'iq_rx', 'iq_tx', 'ce_mean', 'tx_signal')


# ---- variables (tf.get_variable inside tf.layers) -----------------------------------------------------------------
def _variable(g, name, shape, init, trainable=True):
    """VariableV2 + initializer + Assign + read Identity, at the ROOT scope (layer variables live in the variable
    scope, not in the enclosing name scope: `fft_like/conv3d/kernel`, not `receiver/fft_like/...`)."""
    cls = tg.a_strs(['loc:@' + name])
    with g.absolute_scope(''):
        with g.absolute_scope(name + '/Initializer/'):
            if init == 'zeros':
                if len(shape) <= 1 and int(np.prod(shape)) < 1000:
                    iv = tg.const(g, 0.0, FLOAT, 'zeros', splat_shape=list(shape))
                    g.by_name[iv.node].attr['_class'].CopyFrom(cls)
                else:
                    with g.name_scope('zeros'):
                        sh = tg.const(g, list(shape), INT32, 'shape_as_tensor')
                        cv = tg.const(g, 0.0, FLOAT, 'Const')
                    for t in (sh, cv):
                        g.by_name[t.node].attr['_class'].CopyFrom(cls)
                    iv = g.add('Fill', None, [sh, cv], {'T': tg.a_type(FLOAT), 'index_type': tg.a_type(INT32), '_class': cls},
                               [(FLOAT, list(shape))], full_name=name + '/Initializer/zeros')
            else:   # glorot_uniform: random_uniform(shape, -limit, limit)
                limit = float(init)
                with g.name_scope('random_uniform'):
                    sh = tg.const(g, list(shape), INT32, 'shape')
                    mn = tg.const(g, np.float32(-limit), FLOAT, 'min')
                    mx = tg.const(g, np.float32(limit), FLOAT, 'max')
                    ru = g.add('RandomUniform', 'RandomUniform', [sh],
                               {'T': tg.a_type(INT32), 'dtype': tg.a_type(FLOAT), 'seed': tg.a_i(0), 'seed2': tg.a_i(0), '_class': cls},
                               [(FLOAT, list(shape))])
                    sub = g.add('Sub', 'sub', [mx, mn], {'T': tg.a_type(FLOAT), '_class': cls}, [(FLOAT, [])])
                    mul = g.add('Mul', 'mul', [ru, sub], {'T': tg.a_type(FLOAT), '_class': cls}, [(FLOAT, list(shape))])
                for t in (sh, mn, mx):
                    g.by_name[t.node].attr['_class'].CopyFrom(cls)
                iv = g.add('Add', None, [mul, mn], {'T': tg.a_type(FLOAT), '_class': cls}, [(FLOAT, list(shape))],
                           full_name=name + '/Initializer/random_uniform')
        var = g.add('VariableV2', None, [], {'shape': tg.a_shape(shape), 'dtype': tg.a_type(FLOAT), 'container': tg.a_s(''),
                                             'shared_name': tg.a_s(''), '_class': cls}, [(FLOAT, list(shape))], full_name=name)
        g.add('Assign', None, [var, iv], {'T': tg.a_type(FLOAT), 'validate_shape': tg.a_b(True), 'use_locking': tg.a_b(True),
                                          '_class': cls}, [(FLOAT, list(shape))], full_name=name + '/Assign')
        rd = g.add('Identity', None, [var], {'T': tg.a_type(FLOAT), '_class': cls}, [(FLOAT, list(shape))], full_name=name + '/read')
    g.variables.append((var, iv, trainable))
    return rd


def _glorot_limit(shape):
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    return np.sqrt(6.0 / (fan_in + fan_out))


REG_STYLE = {'style': 'contrib', 'scale': 1.0}


def _l2_regularizer(g, layer_scope, var_tail, read):
    """The kernel / bias regulariser of a tf.layers.dense, emitted under `<name scope>/<layer>/<kernel|bias>/Regularizer/`.
    'contrib' (the shipped v1 graphs): tf.contrib.layers.l2_regularizer(scale) = scale * L2Loss(w).
    'keras' (dev/py/model.py:1270-1286): tf.keras.regularizers.l2(0.01) = 0. + 0.01 * reduce_sum(square(w))."""
    cls = tg.a_strs(['loc:@' + read.node[:-len('/read')]])         # colocated with the variable (ops.colocate_with in get_variable)
    scale = REG_STYLE['scale']
    if REG_STYLE['style'] == 'keras':
        with g.absolute_scope(layer_scope + var_tail + '/Regularizer/'):
            sq = g.add('Square', 'Square', [read], {'T': tg.a_type(FLOAT), '_class': cls}, [(FLOAT, read.shape)])
            ax = tg.const(g, list(range(len(read.shape))), INT32, 'Const')
            g.by_name[ax.node].attr['_class'].CopyFrom(cls)
            sm = tg._reduce(g, 'Sum', sq, ax, list(range(len(read.shape))), False, g.unique('Sum'))
            g.by_name[sm.node].attr['_class'].CopyFrom(cls)
            ml = tg.binary(g, 'Mul', np.float32(scale), sm, 'mul')
            out = tg.binary(g, 'Add', np.float32(0.0), ml, 'add')
            for t in (ml, out):
                g.by_name[t.node].attr['_class'].CopyFrom(cls)
                g.by_name[t.node + '/x'].attr['_class'].CopyFrom(cls)
        g.reg_losses.append(out)
        return out
    with g.absolute_scope(layer_scope + var_tail + '/Regularizer/'):
        with g.name_scope('l2_regularizer') as sc:
            s = tg.const(g, np.float32(scale), FLOAT, 'scale')
            g.by_name[s.node].attr['_class'].CopyFrom(cls)
            l2 = g.add('L2Loss', 'L2Loss', [read], {'T': tg.a_type(FLOAT), '_class': cls}, [(FLOAT, [])])
        out = g.add('Mul', None, [s, l2], {'T': tg.a_type(FLOAT), '_class': cls}, [(FLOAT, [])], full_name=sc[:-1])
    g.reg_losses.append(out)
    return out


def _leaky_relu(g, x, name='LeakyRelu'):
    """tf.nn.leaky_relu (TF 1.10): Const alpha, Mul, Maximum -- all under `<name>/`, the Maximum being `<name>` itself."""
    with g.name_scope(name) as sc:
        al = tg.const(g, np.float32(0.2), FLOAT, 'alpha')
        mul = g.add('Mul', 'mul', [al, x], {'T': tg.a_type(FLOAT)}, [(FLOAT, x.shape)])
    return g.add('Maximum', None, [mul, x], {'T': tg.a_type(FLOAT)}, [(FLOAT, x.shape)], full_name=sc[:-1])


def _moments_bn(g, x, eps):
    """tf.nn.moments(x, [0]) + tf.nn.batch_normalization(x, mean, var, None, None, eps) (scopes `moments`, `batchnorm`)."""
    with g.name_scope('moments'):
        mean = tg.reduce_op(g, 'Mean', x, [0], 'mean', keep_dims=True)
        sg = tg.unary(g, 'StopGradient', mean, 'StopGradient')
        sd = g.add('SquaredDifference', 'SquaredDifference', [x, sg], {'T': tg.a_type(FLOAT)}, [(FLOAT, x.shape)])
        var = tg.reduce_op(g, 'Mean', sd, [0], 'variance', keep_dims=True)
        sq_m = g.add('Squeeze', 'Squeeze', [mean], {'T': tg.a_type(FLOAT), 'squeeze_dims': tg.a_ints([0])}, [(FLOAT, x.shape[1:])])
        sq_v = g.add('Squeeze', 'Squeeze_1', [var], {'T': tg.a_type(FLOAT), 'squeeze_dims': tg.a_ints([0])}, [(FLOAT, x.shape[1:])])
    with g.name_scope('batchnorm'):
        add = tg.binary(g, 'Add', sq_v, np.float32(eps), 'add')
        rs = tg.unary(g, 'Rsqrt', add, 'Rsqrt')
        mul = tg.binary(g, 'Mul', x, rs, 'mul')
        neg = tg.unary(g, 'Neg', sq_m, 'Neg')
        mul1 = tg.binary(g, 'Mul', neg, rs, 'mul_1')
        out = tg.binary(g, 'Add', mul, mul1, 'add_1')
    return out


def _power(g, x):
    """tf.reduce_mean(tf.square(x[:,:,:,0]) + tf.square(x[:,:,:,1])) inside the current scope."""
    s0 = tg.strided_slice(g, x, (slice(None), slice(None), slice(None), 0))
    q0 = tg.unary(g, 'Square', s0, 'Square')
    s1 = tg.strided_slice(g, x, (slice(None), slice(None), slice(None), 1))
    q1 = tg.unary(g, 'Square', s1, 'Square')
    ad = tg.binary(g, 'Add', q0, q1, 'add')
    ax = tg.const(g, [0, 1, 2], INT32, 'Const')
    return tg._reduce(g, 'Mean', ad, ax, [0, 1, 2], False, g.unique('Mean'))


def _conv2d_complex(g, x, var_scope, kernel_shape, padding):
    """layers_conv2d_complex (dev/py/complex.py:140-196) on x [B,L,W,C,2] inside the CURRENT name scope: transpose,
    tf.layers.conv3d (variables under `var_scope`), reshape to [...,4,F], the two subtractions, concat, transpose."""
    B, L, W, C, _ = x.shape
    kl, kw, _, _, F2 = kernel_shape
    F = F2 // 2
    conv_in = tg.transpose(g, x, [0, 1, 2, 4, 3])                                        # :168
    k = _variable(g, var_scope + '/kernel', kernel_shape, _glorot_limit(kernel_shape))
    b = _variable(g, var_scope + '/bias', [F2], 'zeros')
    layer = g._scope + var_scope.split('/')[-1]
    Lo, Wo = (L, W) if padding == 'SAME' else (L - kl + 1, W - kw + 1)
    with g.name_scope(var_scope.split('/')[-1]):
        conv = g.add('Conv3D', 'Conv3D', [conv_in, k], {'T': tg.a_type(FLOAT), 'strides': tg.a_ints([1] * 5), 'padding': tg.a_s(padding),
                                                         'dilations': tg.a_ints([1] * 5), 'data_format': tg.a_s('NDHWC')},
                     [(FLOAT, [B, Lo, Wo, 2, F2])])
        conv = g.add('BiasAdd', 'BiasAdd', [conv, b], {'T': tg.a_type(FLOAT), 'data_format': tg.a_s('NHWC')}, [(FLOAT, conv.shape)])
    del layer
    c4 = tg.reshape(g, conv, [-1, Lo, Wo, 4, F])                                          # :185
    c4.shape[0] = None
    sl = lambda i: tg.strided_slice(g, c4, (slice(None), slice(None), slice(None), i, slice(None)))    # noqa: E731
    re = tg.binary(g, 'Sub', sl(0), sl(3), 'sub')                                         # :187
    re = tg.reshape(g, re, [-1, Lo, Wo, 1, F])
    re.shape[0] = None
    im = tg.binary(g, 'Sub', sl(1), sl(2), 'sub')                                         # :188
    im = tg.reshape(g, im, [-1, Lo, Wo, 1, F])
    im.shape[0] = None
    cat = tg.concat(g, [re, im], 3)                                                       # :191
    return tg.transpose(g, cat, [0, 1, 2, 4, 3])                                          # :192


def _dense_matmul(g, x, var_scope, units, regularize=True):
    """tf.layers.dense on a rank-2 input: MatMul + BiasAdd under `<scope>/<layer>/`."""
    n_in = x.shape[-1]
    k = _variable(g, var_scope + '/kernel', [n_in, units], _glorot_limit([n_in, units]))
    layer = var_scope.split('/')[-1]
    scope0 = g._scope
    if regularize:
        _l2_regularizer(g, scope0 + layer + '/', 'kernel', k)
    b = _variable(g, var_scope + '/bias', [units], 'zeros')
    if regularize:
        _l2_regularizer(g, scope0 + layer + '/', 'bias', b)
    with g.name_scope(layer):
        mm = g.add('MatMul', 'MatMul', [x, k], {'T': tg.a_type(FLOAT), 'transpose_a': tg.a_b(False), 'transpose_b': tg.a_b(False)},
                   [(FLOAT, [x.shape[0], units])])
        out = g.add('BiasAdd', 'BiasAdd', [mm, b], {'T': tg.a_type(FLOAT), 'data_format': tg.a_s('NHWC')}, [(FLOAT, mm.shape)])
    return out


def _dense_tensordot(g, x, var_scope, units, activation_leaky):
    """tf.layers.dense on a rank-4 input (TF 1.10 `standard_ops.tensordot(inputs, kernel, [[rank-1],[0]])`): the whole
    `Tensordot/` sub-graph with its shape arithmetic, then BiasAdd [+ LeakyRelu] under `<scope>/<layer>/`."""
    r = x.rank
    n_in = x.shape[-1]
    k = _variable(g, var_scope + '/kernel', [n_in, units], _glorot_limit([n_in, units]))
    layer = var_scope.split('/')[-1]
    scope0 = g._scope
    _l2_regularizer(g, scope0 + layer + '/', 'kernel', k)
    b = _variable(g, var_scope + '/bias', [units], 'zeros')
    _l2_regularizer(g, scope0 + layer + '/', 'bias', b)
    I = INT32
    with g.name_scope(layer) as lsc:
        with g.name_scope('Tensordot') as sc:
            # _tensordot_axes / _tensordot_reshape(a, axes=[r-1]) with a partially known shape (dynamic branch)
            with g.name_scope('range'):
                r0 = tg.const(g, 0, I, 'start')
            rank = tg.const(g, r, I, 'Rank')
            with g.absolute_scope(sc + 'range/'):
                rd = tg.const(g, 1, I, 'delta')
            rng = g.add('Range', None, [r0, rank, rd], {'Tidx': tg.a_type(I)}, [(I, [r])], full_name=sc + 'range')
            axes = tg.const(g, [r - 1], I, 'axes')
            ge = tg.binary(g, 'GreaterEqual', axes, 0, 'GreaterEqual', out_dtype=BOOL)
            c0 = tg.cast(g, ge, I, 'Cast')
            m0 = tg.binary(g, 'Mul', c0, axes, 'mul')
            ls = tg.binary(g, 'Less', axes, 0, 'Less', out_dtype=BOOL)
            c1 = tg.cast(g, ls, I, 'Cast_1')
            ad = tg.binary(g, 'Add', axes, rank, 'add')
            m1 = tg.binary(g, 'Mul', c1, ad, 'mul_1')
            ax2 = tg.binary(g, 'Add', m0, m1, 'add_1')
            free, _ = g.add('ListDiff', 'ListDiff', [rng, ax2], {'T': tg.a_type(I), 'out_idx': tg.a_type(I)}, [(I, [None]), (I, [None])])
            perm = tg.concat(g, [free, ax2], 0, 'concat_1')
            perm.shape = [None]
            tr = g.add('Transpose', 'transpose', [x, perm], {'T': tg.a_type(FLOAT), 'Tperm': tg.a_type(I)}, [(FLOAT, list(x.shape))])
            shp = tg.shape_of(g, x)
            with g.name_scope('GatherV2') as gsc:
                ga = tg.const(g, 0, I, 'axis')
            fd = g.add('GatherV2', None, [shp, free, ga], {'Tparams': tg.a_type(I), 'Tindices': tg.a_type(I), 'Taxis': tg.a_type(I)},
                       [(I, [None])], full_name=gsc[:-1])
            k0 = tg.const(g, [0], I, 'Const')
            p0 = tg._reduce(g, 'Prod', fd, k0, [0], False, g.unique('Prod'))
            with g.name_scope('GatherV2_1') as gsc:
                ga1 = tg.const(g, 0, I, 'axis')
            ad_ = g.add('GatherV2', None, [shp, ax2, ga1], {'Tparams': tg.a_type(I), 'Tindices': tg.a_type(I), 'Taxis': tg.a_type(I)},
                        [(I, [1])], full_name=gsc[:-1])
            k1 = tg.const(g, [0], I, 'Const')
            p1 = tg._reduce(g, 'Prod', ad_, k1, [0], False, g.unique('Prod'))
            stk = g.add('Pack', 'stack', [p0, p1], {'N': tg.a_i(2), 'T': tg.a_type(I), 'axis': tg.a_i(0)}, [(I, [2])])
            a2 = g.add('Reshape', 'Reshape', [tr, stk], {'T': tg.a_type(FLOAT), 'Tshape': tg.a_type(I)}, [(FLOAT, [None, None])])
            # kernel side: static shape -> constants
            kt = tg.transpose(g, k, [0, 1], 'transpose_1')
            k2 = tg.reshape(g, kt, [n_in, units], 'Reshape_1')
            mm = g.add('MatMul', 'MatMul', [a2, k2], {'T': tg.a_type(FLOAT), 'transpose_a': tg.a_b(False), 'transpose_b': tg.a_b(False)},
                       [(FLOAT, [None, units])])
            kc = tg.const(g, [units], I, 'Const')
            osh = tg.concat(g, [fd, kc], 0, 'concat_2')
            osh.shape = [None]
        td = g.add('Reshape', None, [mm, osh], {'T': tg.a_type(FLOAT), 'Tshape': tg.a_type(I)}, [(FLOAT, x.shape[:-1] + [units])],
                   full_name=sc[:-1])
        out = g.add('BiasAdd', 'BiasAdd', [td, b], {'T': tg.a_type(FLOAT), 'data_format': tg.a_s('NHWC')}, [(FLOAT, td.shape)])
        if activation_leaky:
            out = _leaky_relu(g, out)
    del lsc
    return out


def _conv2d_1x1(g, x, var_scope, filters):
    cin = x.shape[-1]
    shape = [1, 1, cin, filters]
    k = _variable(g, var_scope + '/kernel', shape, _glorot_limit(shape))
    b = _variable(g, var_scope + '/bias', [filters], 'zeros')
    with g.name_scope(var_scope.split('/')[-1]):
        cv = g.add('Conv2D', 'Conv2D', [x, k], {'T': tg.a_type(FLOAT), 'strides': tg.a_ints([1, 1, 1, 1]), 'padding': tg.a_s('SAME'),
                                                 'dilations': tg.a_ints([1, 1, 1, 1]), 'data_format': tg.a_s('NHWC'),
                                                 'use_cudnn_on_gpu': tg.a_b(True)}, [(FLOAT, x.shape[:-1] + [filters])])
        return g.add('BiasAdd', 'BiasAdd', [cv, b], {'T': tg.a_type(FLOAT), 'data_format': tg.a_s('NHWC')}, [(FLOAT, cv.shape)])


def _softmax_last(g, x):
    """tf.nn.softmax(x) for rank > 2 (TF 1.10 `_softmax`: flatten to 2-D with dynamic shapes, Softmax, reshape back)."""
    I = INT32
    r = x.rank
    v0 = tg.const(g, [-1], I, 'concat/values_0') if False else None
    del v0
    with g.name_scope('concat'):
        v0 = tg.const(g, [-1], I, 'values_0')
    shp1 = tg.shape_of(g, x, 'Shape_1') if False else None
    del shp1
    # creation order in TF: Rank, Shape, Rank_1, Shape_1, Sub, Slice ...; only names matter here
    rank = tg.const(g, r, I, 'Rank')
    shp = tg.shape_of(g, x, 'Shape')
    rank1 = tg.const(g, r, I, 'Rank_1') if False else rank
    shp1 = tg.shape_of(g, x, 'Shape_1')
    sub = tg.binary(g, 'Sub', rank1, 1, 'Sub')
    with g.name_scope('Slice') as ssc:
        bg = g.add('Pack', 'begin', [sub], {'N': tg.a_i(1), 'T': tg.a_type(I), 'axis': tg.a_i(0)}, [(I, [1])])
        sz = tg.const(g, [1], I, 'size')
    sl = g.add('Slice', None, [shp1, bg, sz], {'T': tg.a_type(I), 'Index': tg.a_type(I)}, [(I, [1])], full_name=ssc[:-1])
    with g.absolute_scope(g._scope + 'concat/'):
        ax = tg.const(g, 0, I, 'axis')
    cc = g.add('ConcatV2', None, [v0, sl, ax], {'N': tg.a_i(2), 'T': tg.a_type(I), 'Tidx': tg.a_type(I)}, [(I, [2])],
               full_name=g._scope + 'concat')
    flat = tg.reshape_dyn(g, x, cc, [None, None], 'Reshape')
    sm = g.add('Softmax', 'Softmax', [flat], {'T': tg.a_type(FLOAT)}, [(FLOAT, [None, None])])
    return tg.reshape_dyn(g, sm, shp, x.shape, 'Reshape')


def _xent(g, logits, labels):
    """tf.nn.softmax_cross_entropy_with_logits[_v2](labels, logits) on rank-2 tensors (TF 1.10 graph)."""
    I = INT32
    with g.name_scope('softmax_cross_entropy_with_logits') as sc:
        rank = tg.const(g, 2, I, 'Rank')
        shp = tg.shape_of(g, logits, 'Shape')
        rank1 = tg.const(g, 2, I, 'Rank_1')
        shp1 = tg.shape_of(g, logits, 'Shape_1')
        sub = tg.binary(g, 'Sub', rank1, 1, 'Sub')
        with g.name_scope('Slice') as ssc:
            bg = g.add('Pack', 'begin', [sub], {'N': tg.a_i(1), 'T': tg.a_type(I), 'axis': tg.a_i(0)}, [(I, [1])])
            sz = tg.const(g, [1], I, 'size')
        sl = g.add('Slice', None, [shp1, bg, sz], {'T': tg.a_type(I), 'Index': tg.a_type(I)}, [(I, [1])], full_name=ssc[:-1])
        with g.name_scope('concat') as csc:
            v0 = tg.const(g, [-1], I, 'values_0')
            ax = tg.const(g, 0, I, 'axis')
        cc = g.add('ConcatV2', None, [v0, sl, ax], {'N': tg.a_i(2), 'T': tg.a_type(I), 'Tidx': tg.a_type(I)}, [(I, [2])], full_name=csc[:-1])
        lf = tg.reshape_dyn(g, logits, cc, [None, None], 'Reshape')
        rank2 = tg.const(g, 2, I, 'Rank_2')
        shp2 = tg.shape_of(g, labels, 'Shape_2')
        sub1 = tg.binary(g, 'Sub', rank2, 1, 'Sub_1')
        with g.name_scope('Slice_1') as ssc:
            bg1 = g.add('Pack', 'begin', [sub1], {'N': tg.a_i(1), 'T': tg.a_type(I), 'axis': tg.a_i(0)}, [(I, [1])])
            sz1 = tg.const(g, [1], I, 'size')
        sl1 = g.add('Slice', None, [shp2, bg1, sz1], {'T': tg.a_type(I), 'Index': tg.a_type(I)}, [(I, [1])], full_name=ssc[:-1])
        with g.name_scope('concat_1') as csc:
            v1 = tg.const(g, [-1], I, 'values_0')
            ax1 = tg.const(g, 0, I, 'axis')
        cc1 = g.add('ConcatV2', None, [v1, sl1, ax1], {'N': tg.a_i(2), 'T': tg.a_type(I), 'Tidx': tg.a_type(I)}, [(I, [2])], full_name=csc[:-1])
        yf = tg.reshape_dyn(g, labels, cc1, [None, None], 'Reshape_1')
        loss, _ = g.add('SoftmaxCrossEntropyWithLogits', None, [lf, yf], {'T': tg.a_type(FLOAT)}, [(FLOAT, [None]), (FLOAT, [None, None])],
                        full_name=sc[:-1])
        sub2 = tg.binary(g, 'Sub', rank, 1, 'Sub_2')
        with g.name_scope('Slice_2') as ssc:
            bg2 = tg.const(g, [0], I, 'begin')
            sz2 = g.add('Pack', 'size', [sub2], {'N': tg.a_i(1), 'T': tg.a_type(I), 'axis': tg.a_i(0)}, [(I, [1])])
        sl2 = g.add('Slice', None, [shp, bg2, sz2], {'T': tg.a_type(I), 'Index': tg.a_type(I)}, [(I, [1])], full_name=ssc[:-1])
        return tg.reshape_dyn(g, loss, sl2, [None], 'Reshape_2')


def _assert_non_negative(g, x, scope_name, message):
    """check_ops.assert_non_negative(x) as TF 1.10 lays it out (assert_less_equal(0, x) with its cond-guarded Assert)."""
    L = INT64
    with g.name_scope(scope_name):
        zero = tg.const(g, 0, L, 'Const')
        with g.name_scope('assert_less_equal'):
            le = g.add('LessEqual', 'LessEqual', [zero, x], {'T': tg.a_type(L)}, [(BOOL, x.shape)])
            ax = tg.const(g, [0], INT32, 'Const')
            al = g.add('All', 'All', [le, ax], {'Tidx': tg.a_type(INT32), 'keep_dims': tg.a_b(False)}, [(BOOL, [])])
            with g.name_scope('Assert'):
                with g.name_scope('AssertGuard') as ag:
                    sw_f, sw_t = g.add('Switch', 'Switch', [al, al], {'T': tg.a_type(BOOL)}, [(BOOL, []), (BOOL, [])])
                    st = g.add('Identity', 'switch_t', [sw_t], {'T': tg.a_type(BOOL)}, [(BOOL, [])])
                    sf = g.add('Identity', 'switch_f', [sw_f], {'T': tg.a_type(BOOL)}, [(BOOL, [])])
                    pid = g.add('Identity', 'pred_id', [al], {'T': tg.a_type(BOOL)}, [(BOOL, [])])
                    noop = g.add('NoOp', 'NoOp', [], {}, [], control=[st])
                    cd = g.add('Identity', 'control_dependency', [st], {'T': tg.a_type(BOOL), '_class': tg.a_strs(['loc:@' + st.node])},
                               [(BOOL, [])], control=[ag + 'NoOp'])
                    del noop
                    with g.name_scope('Assert') as asc:
                        d0 = g.add('Const', 'data_0', [], {'value': tg.a_tensor(message, STRING)[0],
                                                            'dtype': tg.a_type(STRING)}, [(STRING, [])], control=[sf])
                        d1 = g.add('Const', 'data_1', [], {'value': tg.a_tensor('Condition x >= 0 did not hold element-wise:', STRING)[0],
                                                            'dtype': tg.a_type(STRING)}, [(STRING, [])], control=[sf])
                        d2 = g.add('Const', 'data_2', [], {'value': tg.a_tensor('x (%s:0) = ' % x.node, STRING)[0],
                                                            'dtype': tg.a_type(STRING)}, [(STRING, [])], control=[sf])
                        sw0, _ = g.add('Switch', 'Switch', [al, pid], {'T': tg.a_type(BOOL), '_class': tg.a_strs(['loc:@' + al.node])},
                                       [(BOOL, []), (BOOL, [])])
                        sw1, _ = g.add('Switch', 'Switch_1', [x, pid], {'T': tg.a_type(L), '_class': tg.a_strs(['loc:@' + x.node])},
                                       [(L, x.shape), (L, x.shape)])
                    g.add('Assert', None, [sw0, d0, d1, d2, sw1], {'T': tg.a_types([STRING, STRING, STRING, L]), 'summarize': tg.a_i(3)}, [],
                          full_name=asc[:-1])
                    cd1 = g.add('Identity', 'control_dependency_1', [sf], {'T': tg.a_type(BOOL), '_class': tg.a_strs(['loc:@' + sf.node])},
                                [(BOOL, [])], control=[asc[:-1]])
                    mg, _ = g.add('Merge', 'Merge', [cd1, cd], {'N': tg.a_i(2), 'T': tg.a_type(BOOL)}, [(BOOL, []), (INT32, [])])
    return mg


def _confusion_matrix(g, labels, preds):
    """tf.confusion_matrix(labels, predictions) (TF 1.10 confusion_matrix.py) -> int32 [n, n]."""
    L = INT64
    with g.name_scope('confusion_matrix'):
        p = tg.cast(g, preds, L, 'Cast')
        l = tg.cast(g, labels, L, 'Cast_1')
        m_l = _assert_non_negative(g, l, 'assert_non_negative', '`labels` contains negative values')
        l = g.add('Identity', 'control_dependency', [l], {'T': tg.a_type(L), '_class': tg.a_strs(['loc:@' + l.node])}, [(L, l.shape)],
                  control=[m_l])
        m_p = _assert_non_negative(g, p, 'assert_non_negative', '`predictions` contains negative values')
        p = g.add('Identity', 'control_dependency', [p], {'T': tg.a_type(L), '_class': tg.a_strs(['loc:@' + p.node])}, [(L, p.shape)],
                  control=[m_p])
        c0 = tg.const(g, [0], INT32, 'Const')
        mx_p = tg._reduce(g, 'Max', p, c0, [0], False, g.unique('Max'))
        c1 = tg.const(g, [0], INT32, 'Const')
        mx_l = tg._reduce(g, 'Max', l, c1, [0], False, g.unique('Max'))
        mx = g.add('Maximum', 'Maximum', [mx_p, mx_l], {'T': tg.a_type(L)}, [(L, [])])
        n = tg.binary(g, 'Add', mx, np.int64(1), 'add')
        shape = g.add('Pack', 'stack', [n, n], {'N': tg.a_i(2), 'T': tg.a_type(L), 'axis': tg.a_i(0)}, [(L, [2])])
        idx = g.add('Pack', 'stack_1', [l, p], {'N': tg.a_i(2), 'T': tg.a_type(L), 'axis': tg.a_i(1)}, [(L, [None, 2])])
        with g.name_scope('ones_like') as osc:
            osh = tg.shape_of(g, p, 'Shape')
            one = tg.const(g, 1, INT32, 'Const')
        ones = g.add('Fill', None, [osh, one], {'T': tg.a_type(INT32), 'index_type': tg.a_type(INT32)}, [(INT32, [None])], full_name=osc[:-1])
        s32 = tg.cast(g, shape, INT32, 'ToInt32')
        with g.name_scope('zeros') as zsc:
            z0 = tg.const(g, 0, INT32, 'Const')
        zeros = g.add('Fill', None, [s32, z0], {'T': tg.a_type(INT32), 'index_type': tg.a_type(INT32)}, [(INT32, [None, None])], full_name=zsc[:-1])
        return g.add('SparseTensorDenseAdd', 'SparseTensorDenseAdd', [idx, ones, shape, zeros],
                     {'T': tg.a_type(INT32), 'Tindices': tg.a_type(L)}, [(INT32, [None, None])])


def build_receiver_graph(nbits, nsymbol, n_time, n_fft, n_data, nfilter=64, cp=True, head='dev', producer=26):
    """The training graph of ofdmreceiver_np.py up to (and including) `saver = tf.train.Saver()`, without the optimizer.
    n_time = samples per OFDM symbol incl. CP (T), n_fft = K, n_data = frame_size D (dev) or data carriers per symbol x
    symbols (v1: 8 x 46)."""
    g = tg.Graph(producer=producer)
    REG_STYLE.update(style='contrib', scale=1.0) if head == 'v1' else REG_STYLE.update(style='keras', scale=0.01)
    S, T, K, F = nsymbol, n_time, n_fft, nfilter
    Tin = T if cp else K
    v1 = head == 'v1'
    y = tg.placeholder(g, INT32, [None, n_data // (S if v1 else 1), nbits] if not v1 else [None, n_data // S, nbits], 'bits_in')
    x = tg.placeholder(g, FLOAT, [None, S, T, 2], 'tx_ofdm')
    with g.name_scope('transmitter'):                                                     # ofdmreceiver_np.py:125-131
        bn = _moments_bn(g, x, 1e-9)
        iq_tx_re = tg.binary(g, 'RealDiv', bn, np.float32(np.sqrt(2.0)), 'div')
        with g.name_scope('clip_by_norm') as csc:                                         # tf.clip_by_norm(x, 8, axes=[-1])
            sq = tg.binary(g, 'Mul', iq_tx_re, iq_tx_re, 'mul')
            l2 = tg.reduce_op(g, 'Sum', sq, [-1], 'Sum', keep_dims=True)
            nrm = tg.unary(g, 'Sqrt', l2, 'Sqrt')
            inter = tg.binary(g, 'Mul', iq_tx_re, np.float32(8.0), 'mul_1')
            mx = tg.binary(g, 'Maximum', nrm, np.float32(8.0), 'Maximum')
            td = tg.binary(g, 'RealDiv', inter, mx, 'truediv')
        iq_layer = g.add('Identity', None, [td], {'T': tg.a_type(FLOAT)}, [(FLOAT, td.shape)], full_name=csc[:-1])
        power_tx = _power(g, iq_layer)
    snr = tg.placeholder(g, FLOAT, [None, 1], 'SNR')
    with g.name_scope('channel'):                                                         # radio.AWGN_channel, bypassed (:134-138)
        in_real = tg.strided_slice(g, iq_layer, (slice(None), slice(None), slice(None), (0, 1)))
        bn2 = _moments_bn(g, iq_layer, 1e-8)
        inputs = tg.binary(g, 'RealDiv', bn2, np.float32(np.sqrt(2.0)), 'div')
        neg = tg.unary(g, 'Neg', snr, 'Neg')
        d1 = tg.binary(g, 'RealDiv', neg, np.float32(20.0), 'div_1')
        pw = tg.binary(g, 'Pow', np.float32(10.0), d1, 'Pow')
        level = tg.binary(g, 'Mul', np.float32(np.sqrt(.5)), pw, 'mul')
        shp = tg.shape_of(g, in_real, 'Shape')
        with g.name_scope('random_uniform') as rsc:
            mn = tg.const(g, np.float32(0.0), FLOAT, 'min')
            mxv = tg.const(g, np.float32(2 * np.pi), FLOAT, 'max')
            ru = g.add('RandomUniform', 'RandomUniform', [shp], {'T': tg.a_type(INT32), 'dtype': tg.a_type(FLOAT), 'seed': tg.a_i(0),
                                                                  'seed2': tg.a_i(0)}, [(FLOAT, in_real.shape)])
            sb = g.add('Sub', 'sub', [mxv, mn], {'T': tg.a_type(FLOAT)}, [(FLOAT, [])])
            ml = g.add('Mul', 'mul', [ru, sb], {'T': tg.a_type(FLOAT)}, [(FLOAT, in_real.shape)])
        phase = g.add('Add', None, [ml, mn], {'T': tg.a_type(FLOAT)}, [(FLOAT, in_real.shape)], full_name=rsc[:-1])
        lv4 = tg.reshape(g, level, [-1, 1, 1, 1])
        lv4.shape[0] = None
        shp1 = tg.shape_of(g, in_real, 'Shape')
        with g.name_scope('random_normal') as nsc:
            mean = tg.const(g, np.float32(0.0), FLOAT, 'mean')
            std = tg.const(g, np.float32(1.0), FLOAT, 'stddev')
            rn = g.add('RandomStandardNormal', 'RandomStandardNormal', [shp1], {'T': tg.a_type(INT32), 'dtype': tg.a_type(FLOAT),
                                                                                 'seed': tg.a_i(0), 'seed2': tg.a_i(0)}, [(FLOAT, in_real.shape)])
            ml = g.add('Mul', 'mul', [rn, std], {'T': tg.a_type(FLOAT)}, [(FLOAT, in_real.shape)])
        rnd = g.add('Add', None, [ml, mean], {'T': tg.a_type(FLOAT)}, [(FLOAT, in_real.shape)], full_name=nsc[:-1])
        amp = tg.binary(g, 'Mul', lv4, rnd, 'Mul')
        amp.shape = in_real.shape
        ab = tg.unary(g, 'Abs', amp, 'Abs')
        sn = tg.unary(g, 'Sin', phase, 'Sin')
        n_re = tg.binary(g, 'Mul', ab, sn, 'Mul')
        ab1 = tg.unary(g, 'Abs', amp, 'Abs')
        cs = tg.unary(g, 'Cos', phase, 'Cos')
        n_im = tg.binary(g, 'Mul', ab1, cs, 'Mul')
        noise = tg.concat(g, [n_re, n_im], -1)
        iq_receiver = tg.binary(g, 'Add', inputs, noise, 'add')
        noise_pwr = _power(g, noise)
    rx_in = iq_tx_re                                                                      # :138 (in-graph AWGN bypassed)
    with g.name_scope('receiver'):                                                        # ofdm_dense_rx, model.py:1222-1292
        src = rx_in
        if not cp:                                                                        # tf.slice(out, [0,0,CP,0], [-1,-1,K,-1]) :1236-1238
            with g.name_scope('Slice') as ssc:
                bg = tg.const(g, [0, 0, T - K, 0], INT32, 'begin')
                sz = tg.const(g, [-1, -1, K, -1], INT32, 'size')
            src = g.add('Slice', None, [rx_in, bg, sz], {'T': tg.a_type(FLOAT), 'Index': tg.a_type(INT32)}, [(FLOAT, [None, S, K, 2])],
                        full_name=ssc[:-1])
        with g.name_scope('fft_like'):
            ci = tg.reshape(g, src, [-1, S, 1, Tin, 2])
            ci.shape[0] = None
            ff = _conv2d_complex(g, ci, 'fft_like/conv3d', [1, Tin, 1, Tin, 2 * F], 'SAME')
            ff = tg.reshape(g, ff, [-1, S, F, 2])
            ff.shape[0] = None
        with g.name_scope('demodulation'):
            flat = tg.reshape(g, ff, [-1, S * F * 2])
            flat.shape[0] = None
            dn = _dense_matmul(g, flat, 'demodulation/dense', 2 * n_data)
            if v1:
                oiq = tg.reshape(g, dn, [-1, S, n_data // S, 2])
            else:
                oiq = tg.reshape(g, dn, [-1, 1, n_data, 2])
            oiq.shape[0] = None
            h = _conv2d_1x1(g, oiq, 'demodulation/conv2d', 1 << nbits)
            if v1:
                h = _conv2d_1x1(g, h, 'demodulation/conv2d_1', 1 << nbits)
            h = _leaky_relu(g, h)
            cat = tg.concat(g, [h, oiq], -1)
            lg = _dense_tensordot(g, cat, 'demodulation/dense_1', 2 * nbits, True)
        lg = tg.reshape(g, lg, [-1, n_data // S if v1 else n_data, nbits, 2])
        lg.shape[0] = None
        outputs = _softmax_last(g, lg)
    iq_tx = tg.cast(g, _rs(g, iq_layer, [-1, 2]), HALF)                                   # :148-149
    iq_rx = tg.cast(g, _rs(g, iq_receiver, [-1, 2]), HALF)
    y_flat = _rs(g, y, [-1])
    with g.name_scope('one_hot') as osc:
        on = tg.const(g, np.float32(1.0), FLOAT, 'on_value')
        off = tg.const(g, np.float32(0.0), FLOAT, 'off_value')
        dp = tg.const(g, 2, INT32, 'depth')
    y1h = g.add('OneHot', None, [y_flat, dp, on, off], {'T': tg.a_type(FLOAT), 'TI': tg.a_type(INT32), 'axis': tg.a_i(-1)},
                [(FLOAT, [None, 2])], full_name=osc[:-1])
    out_sm = _rs(g, outputs, [-1, 2])
    xent = _xent(g, out_sm, y1h)
    ax = tg.const(g, [0], INT32, 'Const')
    ce_mean = tg._reduce(g, 'Mean', xent, ax, [0], False, g.unique('ce_mean'))
    y_index = _rs(g, y, [-1])
    with g.name_scope('ArgMax') as asc:
        dim = tg.const(g, 1, INT32, 'dimension')
    am = g.add('ArgMax', None, [out_sm, dim], {'T': tg.a_type(FLOAT), 'Tidx': tg.a_type(INT32), 'output_type': tg.a_type(INT64)},
               [(INT64, [None])], full_name=asc[:-1])
    out_index = tg.cast(g, am, INT32)
    conf = _confusion_matrix(g, y_index, out_index)
    # ber_tensor (util.py:44-48)
    ax2 = tg.const(g, [0, 1], INT32, 'Const')
    total = tg._reduce(g, 'Sum', conf, ax2, [0, 1], False, g.unique('Sum'))
    e01 = tg.strided_slice(g, conf, (0, 1))
    e10 = tg.strided_slice(g, conf, (1, 0))
    err = tg.binary(g, 'Add', e01, e10, 'add')
    with g.name_scope('BER') as bsc:
        c_e = tg.cast(g, err, tg.types_pb2.DT_DOUBLE, 'Cast')
        c_t = tg.cast(g, total, tg.types_pb2.DT_DOUBLE, 'Cast_1')
    ber_lin = g.add('RealDiv', None, [c_e, c_t], {'T': tg.a_type(tg.types_pb2.DT_DOUBLE)}, [(tg.types_pb2.DT_DOUBLE, [])], full_name=bsc[:-1])
    ber_log = g.add('Log', 'Log', [ber_lin], {'T': tg.a_type(tg.types_pb2.DT_DOUBLE)}, [(tg.types_pb2.DT_DOUBLE, [])])
    berlin = tg.cast(g, ber_lin, FLOAT)
    # total_loss = ce_mean + berlin * REG_COEFF * sum(regularization_losses) + BER_COEFF * cast(ber)       (:162-171)
    t1 = tg.binary(g, 'Mul', berlin, np.float32(0.0001), 'mul')
    acc = None
    for i, r in enumerate(g.reg_losses):                                                  # sum(): 0 + r0 + r1 + ...
        acc = tg.binary(g, 'Add', np.float32(0.0) if acc is None else acc, r, 'add')
    t2 = tg.binary(g, 'Mul', t1, acc, 'mul')
    t3 = tg.binary(g, 'Add', ce_mean, t2, 'add')
    c_b = tg.cast(g, ber_log, FLOAT)
    t4 = tg.binary(g, 'Mul', np.float32(1.0), c_b, 'mul')
    total_loss = tg.binary(g, 'Add', t3, t4, 'add')
    tg.identity(g, iq_layer, 'tx_signal')
    tg.identity(g, power_tx, 'tx_power')
    tg.identity(g, rx_in, 'input')
    tg.identity(g, outputs, 'output')
    tg.identity(g, total_loss, 'cost')
    tg.identity(g, ber_log, 'log_ber')
    tg.identity(g, berlin, 'linear_ber')
    tg.identity(g, conf, 'conf_matrix')
    tg.identity(g, noise_pwr, 'noise_power')
    tg.identity(g, iq_rx, 'iq_rx')
    tg.identity(g, iq_tx, 'iq_tx')
    _variable(g, 'global_step', [], 'zeros', trainable=False)
    return g


def _rs(g, x, shape):
    out = tg.reshape(g, x, shape)
    out.shape = [None] + list(shape[1:])
    return out


# ---- tf.train.Saver(): the save/ sub-graph (V2 format) and the MetaGraphDef around it ---------------------------------
def add_saver(g, extra_names=()):
    """`tf.train.Saver()` over every variable of the graph, names sorted like BaseSaverBuilder does."""
    names = sorted([v.node for v, _, _ in g.variables] + list(extra_names))
    var = {v.node: v for v, _, _ in g.variables}
    with g.absolute_scope(''):
        with g.name_scope('save') as sc:
            fn = g.add('Const', 'Const', [], {'value': tg.a_tensor('model', STRING)[0], 'dtype': tg.a_type(STRING)}, [(STRING, [])])
            with g.name_scope('SaveV2') as ssc:
                tn = g.add('Const', 'tensor_names', [], {'value': tg.a_tensor(list(names), STRING)[0], 'dtype': tg.a_type(STRING)},
                           [(STRING, [len(names)])])
                ss = g.add('Const', 'shape_and_slices', [], {'value': tg.a_tensor([''] * len(names), STRING)[0], 'dtype': tg.a_type(STRING)},
                           [(STRING, [len(names)])])
            g.add('SaveV2', None, [fn, tn, ss] + [var[n] for n in names], {'dtypes': tg.a_types([FLOAT] * len(names))}, [], full_name=ssc[:-1])
            g.add('Identity', 'control_dependency', [fn], {'T': tg.a_type(STRING), '_class': tg.a_strs(['loc:@save/Const'])}, [(STRING, [])],
                  control=[ssc[:-1]])
            with g.name_scope('RestoreV2') as rsc:
                tn2 = g.add('Const', 'tensor_names', [], {'value': tg.a_tensor(list(names), STRING)[0], 'dtype': tg.a_type(STRING)},
                            [(STRING, [len(names)])])
                ss2 = g.add('Const', 'shape_and_slices', [], {'value': tg.a_tensor([''] * len(names), STRING)[0], 'dtype': tg.a_type(STRING)},
                            [(STRING, [len(names)])])
            for t in (tn2, ss2):
                g.by_name[t.node].device = '/device:CPU:0'
            outs = g.add('RestoreV2', None, [fn, tn2, ss2], {'dtypes': tg.a_types([FLOAT] * len(names))}, [(FLOAT, None)] * len(names),
                         full_name=rsc[:-1])
            outs = outs if isinstance(outs, list) else [outs]
            g.by_name[rsc[:-1]].device = '/device:CPU:0'
            assigns = []
            for i, n in enumerate(names):
                a = g.add('Assign', 'Assign', [var[n], outs[i]], {'T': tg.a_type(FLOAT), 'validate_shape': tg.a_b(True), 'use_locking': tg.a_b(True),
                                                                  '_class': tg.a_strs(['loc:@' + n])}, [(FLOAT, var[n].shape)])
                assigns.append(a.node)
            g.add('NoOp', 'restore_all', [], {}, [], control=sorted(assigns))
    sd = saver_pb2.SaverDef(filename_tensor_name='save/Const:0', save_tensor_name='save/control_dependency:0',
                            restore_op_name='save/restore_all', max_to_keep=5, keep_checkpoint_every_n_hours=10000.0,
                            version=saver_pb2.SaverDef.V2)
    return sd


def meta_graph(g, tf_version='1.15.0', git_version='v1.15.0-0-g590d6eef7e'):
    m = meta_graph_pb2.MetaGraphDef()
    sd = add_saver(g)
    with g.absolute_scope(''):
        g.add('NoOp', 'init', [], {}, [], control=sorted(v.node + '/Assign' for v, _, _ in g.variables))
    m.meta_info_def.tensorflow_version = tf_version
    m.meta_info_def.tensorflow_git_version = git_version
    m.graph_def.CopyFrom(g.graph_def())
    m.saver_def.CopyFrom(sd)
    for key, sel in (('variables', lambda t: True), ('trainable_variables', lambda t: t)):
        for v, iv, trainable in g.variables:
            if sel(trainable):
                vd = variable_pb2.VariableDef(variable_name=v.name, initial_value_name=iv.name, initializer_name=v.node + '/Assign',
                                              snapshot_name=v.node + '/read:0', trainable=trainable)
                m.collection_def[key].bytes_list.value.append(vd.SerializeToString())
    for r in g.reg_losses:
        m.collection_def['regularization_losses'].node_list.value.append(r.name)
    return m


def write_meta(path, nbits, ofdmobj, nfilter=64, cp=True, head='dev'):
    """Write `path`.meta for the basic receiver of geometry `ofdmobj` (an ofdm_tx)."""
    g = build_receiver_graph(nbits, ofdmobj.nSymbol, ofdmobj.K + ofdmobj.CP, ofdmobj.K, ofdmobj.frame_size, nfilter, cp, head,
                             producer=134)
    m = meta_graph(g)
    with open(path + '.meta', 'wb') as f:
        f.write(m.SerializeToString())
    return m
