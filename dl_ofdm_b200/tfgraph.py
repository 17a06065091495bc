"""A TensorFlow-1.x GraphDef emitter without TensorFlow: just enough of TF's Python op library to rebuild, node for node,
the inference graph `ofdmreceiver_np.py` builds (dev/py/ofdmreceiver_np.py:121-183) -- names, ops, inputs, attributes and
static shapes as TF 1.10-1.15 would emit them -- so that `tfmeta.write_meta` can put a `.meta` file next to the
`.index` / `.data` bundle and the reference's `load_model_np` (dev/py/model.py:51-72: `tf.train.import_meta_graph` +
`saver.restore`, then `graph.get_tensor_by_name('output:0')` ...) finds the graph it expects.

The emitter follows TF's naming rules (`Graph.unique_name`: `_1`, `_2` suffixes per scope; `name_scope`; the auto-named
`Const` inputs such as `.../Reshape/shape`, `.../strided_slice/stack_1`), and each helper below emits exactly the nodes
the corresponding `tf.*` call emits.  tests/test_host.py compares the result with the graphs the reference ships
(`test_v1/model/*.meta`, written by TF 1.10.1): every node reachable from the named fetches must match in name, op,
inputs and the attributes listed in `CHECKED_ATTRS`.  Protobuf classes come from `tensorboard.compat.proto`.
"""
from __future__ import annotations

import contextlib

import numpy as np
from tensorboard.compat.proto import (attr_value_pb2, graph_pb2, node_def_pb2, tensor_pb2, tensor_shape_pb2,
                                      types_pb2)

FLOAT, INT32, INT64, BOOL, STRING, HALF = (types_pb2.DT_FLOAT, types_pb2.DT_INT32, types_pb2.DT_INT64,
                                           types_pb2.DT_BOOL, types_pb2.DT_STRING, types_pb2.DT_HALF)
_NP = {FLOAT: np.float32, INT32: np.int32, INT64: np.int64, BOOL: np.bool_, HALF: np.float16}


class T:
    """Symbolic tensor: producing node, output index, dtype, static shape (None = unknown dimension)."""

    def __init__(self, node, idx, dtype, shape):
        self.node, self.idx, self.dtype = node, idx, dtype
        self.shape = None if shape is None else list(shape)

    @property
    def ref(self):
        return self.node if self.idx == 0 else '%s:%d' % (self.node, self.idx)

    @property
    def name(self):
        return '%s:%d' % (self.node, self.idx)

    @property
    def rank(self):
        return len(self.shape)


def _shape_proto(shape):
    p = tensor_shape_pb2.TensorShapeProto()
    if shape is None:
        p.unknown_rank = True
        return p
    for d in shape:
        p.dim.add().size = -1 if d is None else int(d)
    return p


def _tensor_proto(value, dtype, splat_shape=None):
    """splat_shape: `tf.constant(scalar, shape=...)` -- the shape with ONE typed value entry (how TF stores zeros(...))."""
    t = tensor_pb2.TensorProto(dtype=dtype)
    if splat_shape is not None:
        t.tensor_shape.CopyFrom(_shape_proto(splat_shape))
        field = {FLOAT: t.float_val, INT32: t.int_val, INT64: t.int64_val, BOOL: t.bool_val}[dtype]
        field.append(np.asarray(value, dtype=_NP[dtype]).item())
        return t, list(splat_shape)
    if dtype == STRING:
        vals = value if isinstance(value, (list, tuple)) else [value]
        shape = [len(vals)] if isinstance(value, (list, tuple)) else []
        t.tensor_shape.CopyFrom(_shape_proto(shape))
        for v in vals:
            t.string_val.append(v if isinstance(v, bytes) else v.encode())
        return t, shape
    a = np.asarray(value, dtype=_NP[dtype])
    t.tensor_shape.CopyFrom(_shape_proto(a.shape))
    flat = a.reshape(-1)
    # TF's make_tensor_proto: scalars / single elements and all-equal arrays go to the typed *_val field (one entry),
    # everything else to tensor_content
    if a.size == 1 or (a.size > 0 and np.all(flat == flat[0]) and a.ndim <= 1 and a.size == 1):
        field = {FLOAT: t.float_val, INT32: t.int_val, INT64: t.int64_val, BOOL: t.bool_val, HALF: t.half_val}[dtype]
        field.append(flat[0].item() if dtype != HALF else int(flat[:1].view(np.uint16)[0]))
    elif a.size > 0:
        t.tensor_content = a.tobytes()
    return t, list(a.shape)


class Graph:
    def __init__(self, producer=26):
        self.nodes = []              # NodeDef, in creation order
        self.by_name = {}
        self._used = {}
        self._scope = ''
        self.producer = producer
        self.variables = []          # (T var, initial value T, trainable)
        self.reg_losses = []

    # ---- naming (tf.Graph.unique_name / name_scope) ---------------------------------------------
    def unique(self, name, mark=True):
        full = self._scope + name
        key = full.lower()
        i = self._used.get(key, 0)
        if mark:
            self._used[key] = i + 1
        if i > 0:
            base = key
            while True:
                cand = '%s_%d' % (full, i)
                if cand.lower() not in self._used:
                    break
                i += 1
            if mark:
                self._used[base] = i + 1
                self._used[cand.lower()] = 1
            full = cand
        return full

    @contextlib.contextmanager
    def name_scope(self, name):
        old = self._scope
        self._scope = self.unique(name) + '/'
        try:
            yield self._scope
        finally:
            self._scope = old

    @contextlib.contextmanager
    def absolute_scope(self, scope):
        """Re-enter an existing scope ('a/b/' -- tf.name_scope('a/b/')) or the root ('')."""
        old = self._scope
        self._scope = scope
        try:
            yield scope
        finally:
            self._scope = old

    # ---- node emission ---------------------------------------------------------------------------
    def add(self, op, name, inputs, attrs, out, full_name=None, control=()):
        """out: list of (dtype, shape) per output.  Returns the first output (or a list when several)."""
        nd = node_def_pb2.NodeDef()
        nd.name = full_name if full_name is not None else self.unique(name)
        nd.op = op
        for i in inputs:
            nd.input.append(i.ref if isinstance(i, T) else i)
        for c in control:
            nd.input.append('^' + (c.node if isinstance(c, T) else c))
        for k, v in attrs.items():
            nd.attr[k].CopyFrom(v)
        if out:
            lst = attr_value_pb2.AttrValue.ListValue()
            for _, shp in out:
                lst.shape.add().CopyFrom(_shape_proto(shp))
            nd.attr['_output_shapes'].CopyFrom(attr_value_pb2.AttrValue(list=lst))
        assert nd.name not in self.by_name, nd.name
        self.nodes.append(nd)
        self.by_name[nd.name] = nd
        outs = [T(nd.name, i, dt, shp) for i, (dt, shp) in enumerate(out)]
        return outs[0] if len(outs) == 1 else outs

    def graph_def(self):
        g = graph_pb2.GraphDef()
        g.versions.producer = self.producer
        g.node.extend(self.nodes)
        return g


# ---- attribute helpers ---------------------------------------------------------------------------
def a_type(t):
    return attr_value_pb2.AttrValue(type=t)


def a_b(v):
    return attr_value_pb2.AttrValue(b=bool(v))


def a_i(v):
    return attr_value_pb2.AttrValue(i=int(v))


def a_s(v):
    return attr_value_pb2.AttrValue(s=v if isinstance(v, bytes) else v.encode())


def a_shape(shape):
    return attr_value_pb2.AttrValue(shape=_shape_proto(shape))


def a_ints(vs):
    return attr_value_pb2.AttrValue(list=attr_value_pb2.AttrValue.ListValue(i=[int(v) for v in vs]))


def a_types(ts):
    return attr_value_pb2.AttrValue(list=attr_value_pb2.AttrValue.ListValue(type=list(ts)))


def a_strs(vs):
    return attr_value_pb2.AttrValue(list=attr_value_pb2.AttrValue.ListValue(s=[v.encode() for v in vs]))


def a_tensor(value, dtype, splat_shape=None):
    t, shape = _tensor_proto(value, dtype, splat_shape)
    return attr_value_pb2.AttrValue(tensor=t), shape


# ---- the subset of tf.* used by the reference graph ----------------------------------------------
def const(g, value, dtype, name='Const', splat_shape=None):
    av, shape = a_tensor(value, dtype, splat_shape)
    return g.add('Const', name, [], {'value': av, 'dtype': a_type(dtype)}, [(dtype, shape)])


def placeholder(g, dtype, shape, name):
    return g.add('Placeholder', name, [], {'dtype': a_type(dtype), 'shape': a_shape(shape)}, [(dtype, shape)])


def _bshape(a, b):
    ra, rb = list(a.shape), list(b.shape)
    n = max(len(ra), len(rb))
    ra, rb = [1] * (n - len(ra)) + ra, [1] * (n - len(rb)) + rb
    out = []
    for x, y in zip(ra, rb):
        if x == 1:
            out.append(y)
        elif y == 1 or x == y:
            out.append(x)
        elif x is None or y is None:
            out.append(x if y is None else y)
        else:
            raise ValueError('shapes %s %s' % (a.shape, b.shape))
    return out


def binary(g, op, a, b, name, out_dtype=None):
    """Add / Sub / Mul / RealDiv / Maximum / ... ; a Python scalar operand becomes the auto-named Const `<name>/x|y`."""
    with g.name_scope(name) as scope:
        if not isinstance(a, T):
            a = const(g, a, b.dtype, 'x')
        if not isinstance(b, T):
            b = const(g, b, a.dtype, 'y')
    attrs = {'T': a_type(a.dtype)}
    return g.add(op, None, [a, b], attrs, [(out_dtype or a.dtype, _bshape(a, b))], full_name=scope[:-1])


def unary(g, op, x, name, **extra):
    return g.add(op, name, [x], dict({'T': a_type(x.dtype)}, **extra), [(x.dtype, x.shape)])


def identity(g, x, name):
    return g.add('Identity', name, [x], {'T': a_type(x.dtype)}, [(x.dtype, x.shape)])


def cast(g, x, dtype, name='Cast'):
    return g.add('Cast', name, [x], {'SrcT': a_type(x.dtype), 'DstT': a_type(dtype)}, [(dtype, x.shape)])


def reshape(g, x, shape, name='Reshape'):
    """tf.reshape(x, [python ints]) -> Const `<name>/shape` + Reshape."""
    with g.name_scope(name) as scope:
        s = const(g, shape, INT32, 'shape')
    known = [d for d in shape if d != -1]
    out = [None if d == -1 else d for d in shape]
    if -1 in shape and x.shape is not None and all(d is not None for d in x.shape):
        out[shape.index(-1)] = int(np.prod(x.shape)) // max(1, int(np.prod(known)))
    return g.add('Reshape', None, [x, s], {'T': a_type(x.dtype), 'Tshape': a_type(INT32)}, [(x.dtype, out)],
                 full_name=scope[:-1])


def reshape_dyn(g, x, shape_t, out_shape, name='Reshape'):
    return g.add('Reshape', name, [x, shape_t], {'T': a_type(x.dtype), 'Tshape': a_type(INT32)}, [(x.dtype, out_shape)])


def transpose(g, x, perm, name='transpose'):
    with g.name_scope(name) as scope:
        p = const(g, perm, INT32, 'perm')
    return g.add('Transpose', None, [x, p], {'T': a_type(x.dtype), 'Tperm': a_type(INT32)},
                 [(x.dtype, [x.shape[i] for i in perm])], full_name=scope[:-1])


def concat(g, xs, axis, name='concat'):
    with g.name_scope(name) as scope:
        ax = const(g, axis, INT32, 'axis')
    r = xs[0].rank
    a = axis % r
    out = list(xs[0].shape)
    out[a] = None if any(x.shape[a] is None for x in xs) else sum(x.shape[a] for x in xs)
    return g.add('ConcatV2', None, list(xs) + [ax], {'N': a_i(len(xs)), 'T': a_type(xs[0].dtype), 'Tidx': a_type(INT32)},
                 [(xs[0].dtype, out)], full_name=scope[:-1])


def strided_slice(g, x, spec, name='strided_slice'):
    """x[spec] with spec a tuple of slice(None) / (begin, end) / int, exactly like Tensor.__getitem__: the three Const inputs
    `<name>/stack`, `/stack_1`, `/stack_2` and the begin / end / shrink masks."""
    begin, end, strides = [], [], []
    bm = em = sm = 0
    out = []
    for i, s in enumerate(spec):
        if isinstance(s, int):
            begin.append(s); end.append(s + 1); strides.append(1)
            sm |= 1 << i
        elif s == slice(None):
            begin.append(0); end.append(0); strides.append(1)
            bm |= 1 << i; em |= 1 << i
            out.append(x.shape[i])
        else:
            b, e = s
            begin.append(0 if b is None else b); end.append(0 if e is None else e); strides.append(1)
            if b is None:
                bm |= 1 << i
            if e is None:
                em |= 1 << i
            lo = 0 if b is None else b
            hi = x.shape[i] if e is None else e
            out.append(None if hi is None else hi - lo)
    out += x.shape[len(spec):]
    with g.name_scope(name) as scope:
        s0 = const(g, begin, INT32, 'stack')
        s1 = const(g, end, INT32, 'stack_1')
        s2 = const(g, strides, INT32, 'stack_2')
    attrs = {'T': a_type(x.dtype), 'Index': a_type(INT32), 'begin_mask': a_i(bm), 'end_mask': a_i(em),
             'ellipsis_mask': a_i(0), 'new_axis_mask': a_i(0), 'shrink_axis_mask': a_i(sm)}
    return g.add('StridedSlice', None, [x, s0, s1, s2], attrs, [(x.dtype, out)], full_name=scope[:-1])


def reduce_op(g, op, x, axes, name, keep_dims=False, axes_name='reduction_indices'):
    """tf.reduce_mean / reduce_sum / reduce_max(x, axes): Const `<name>/reduction_indices` + the reduction node."""
    with g.name_scope(name) as scope:
        ax = const(g, axes, INT32, axes_name)
    return _reduce(g, op, x, ax, axes, keep_dims, scope[:-1])


def _reduce(g, op, x, ax_t, axes, keep_dims, full_name):
    al = [a % x.rank for a in (axes if isinstance(axes, (list, tuple)) else [axes])]
    out = [(1 if i in al else d) for i, d in enumerate(x.shape)] if keep_dims else [d for i, d in enumerate(x.shape) if i not in al]
    return g.add(op, None, [x, ax_t], {'T': a_type(x.dtype), 'Tidx': a_type(INT32), 'keep_dims': a_b(keep_dims)},
                 [(x.dtype, out)], full_name=full_name)


def reduce_all_axes(g, op, x, name):
    """tf.reduce_mean(x) / tf.reduce_sum(x) with axis=None: a root-level `Const` [0..rank) created BEFORE the op's scope."""
    ax = const(g, list(range(x.rank)), INT32, 'Const')
    return _reduce(g, op, x, ax, list(range(x.rank)), False, g.unique(name))


def shape_of(g, x, name='Shape'):
    return g.add('Shape', name, [x], {'T': a_type(x.dtype), 'out_type': a_type(INT32)}, [(INT32, [x.rank])])
