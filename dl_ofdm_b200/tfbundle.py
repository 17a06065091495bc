"""Pure-Python reader / writer for TensorFlow "bundle v2" checkpoints.

The reference saves and restores its models with ``tf.train.Saver``
(reference: dev/py/ofdmreceiver_np.py:192,268-272, dev/py/model.py:51-56) which
produces ``<prefix>.index`` (a LevelDB-style SSTable of ``BundleEntryProto``)
and ``<prefix>.data-00000-of-00001`` (raw little-endian tensor bytes).  The
eight trained v1 receivers under ``test_v1/model/`` use that format.  TensorFlow
itself is not available in this image, so this module decodes / encodes the
format directly; no TensorFlow code is involved.

Only what the DCCN path needs is implemented: uncompressed blocks, a single
data shard, dtypes float32 / int32 / int64.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DT = {1: np.float32, 3: np.int32, 9: np.int64, 2: np.float64}
_DT_INV = {np.dtype(np.float32): 1, np.dtype(np.int32): 3, np.dtype(np.int64): 9,
           np.dtype(np.float64): 2}


# ----------------------------------------------------------------------------
# varint / protobuf helpers
# ----------------------------------------------------------------------------
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """Minimal protobuf wire decoder: field number -> list of raw values."""
    pos = 0
    out: Dict[int, list] = {}
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        out.setdefault(field, []).append(v)
    return out


def _parse_entry(val: bytes):
    p = _parse_proto(val)
    dtype = p.get(1, [0])[0]
    shape: List[int] = []
    if 2 in p:
        sp = _parse_proto(p[2][0])
        for d in sp.get(2, []):
            dp = _parse_proto(d)
            shape.append(dp.get(1, [0])[0])
    offset = p.get(4, [0])[0]
    size = p.get(5, [0])[0]
    return dtype, tuple(shape), offset, size


# ----------------------------------------------------------------------------
# SSTable reading
# ----------------------------------------------------------------------------
def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        val = block[pos:pos + vlen]
        pos += vlen
        yield key, val


def _read_block(buf: bytes, offset: int, size: int) -> bytes:
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError('compressed SSTable blocks are not supported (type %d)' % ctype)
    return buf[offset:offset + size]


def read_index(prefix: str) -> Dict[str, Tuple[int, Tuple[int, ...], int, int]]:
    """name -> (dtype enum, shape, offset, size) for every tensor in ``prefix``.index."""
    buf = open(prefix + '.index', 'rb').read()
    footer = buf[-48:]
    if struct.unpack_from('<Q', footer, 40)[0] != _MAGIC:
        raise ValueError('%s.index: bad SSTable magic' % prefix)
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex offset
    _, pos = _get_varint(footer, pos)          # metaindex size
    ioff, pos = _get_varint(footer, pos)
    isz, pos = _get_varint(footer, pos)
    out = {}
    for _, handle in _block_entries(_read_block(buf, ioff, isz)):
        boff, p = _get_varint(handle, 0)
        bsz, p = _get_varint(handle, p)
        for key, val in _block_entries(_read_block(buf, boff, bsz)):
            if key == b'':
                continue                        # BundleHeaderProto
            out[key.decode()] = _parse_entry(val)
    return out


def read_checkpoint(prefix: str, model_only: bool = True) -> Dict[str, np.ndarray]:
    """Load every tensor of a TF bundle as numpy arrays keyed by variable name.

    ``model_only`` drops optimizer slots (``.../Adam``, ``.../Adam_1``,
    ``beta?_power``) the way ``dccn_set_weight`` would ignore them anyway.
    """
    index = read_index(prefix)
    data = np.memmap(prefix + '.data-00000-of-00001', dtype=np.uint8, mode='r')
    out = {}
    for name, (dt, shape, off, size) in index.items():
        if model_only and (name.endswith('/Adam') or name.endswith('/Adam_1')
                           or name.startswith('beta1_power') or name.startswith('beta2_power')):
            continue
        if dt not in _DT:
            continue
        arr = np.frombuffer(bytes(data[off:off + size]), dtype=_DT[dt]).reshape(shape)
        out[name] = arr.copy()
    return out


# ----------------------------------------------------------------------------
# SSTable writing (enough for tf.train.Saver().restore to read back)
# ----------------------------------------------------------------------------
def _crc32c_table():
    poly = 0x82F63B78
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TBL = _crc32c_table()


def _crc32c_native():
    """dccn_crc32c of libdccn.so (host code, works without a GPU) or None when the library has not been built."""
    try:
        from . import _lib
        return _lib.load().dccn_crc32c
    except Exception:
        return None


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C.  Large buffers go through the library's slicing-by-8 routine (a checkpoint holds 12 MB of tensors and
    is rewritten on every improving epoch); the pure-Python loop below is the small-input / no-library path."""
    if len(data) >= 4096:
        fn = _crc32c_native()
        if fn is not None:
            buf = bytes(data)
            return int(fn(buf, len(buf), crc))
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = _CRC_TBL[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask_crc(c: int) -> int:
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _proto_field(field: int, wt: int, payload) -> bytes:
    key = _put_varint((field << 3) | wt)
    if wt == 0:
        return key + _put_varint(payload)
    if wt == 2:
        return key + _put_varint(len(payload)) + payload
    if wt == 5:
        return key + struct.pack('<I', payload)
    raise ValueError(wt)


def _entry_proto(arr: np.ndarray, offset: int) -> bytes:
    shape = b''.join(_proto_field(2, 2, _proto_field(1, 0, int(d))) for d in arr.shape)
    raw = arr.tobytes()
    out = _proto_field(1, 0, _DT_INV[arr.dtype])
    out += _proto_field(2, 2, shape)
    if offset:
        out += _proto_field(4, 0, offset)
    out += _proto_field(5, 0, len(raw))
    out += _proto_field(6, 5, _mask_crc(crc32c(raw)))
    return out


def _build_block(items: List[Tuple[bytes, bytes]]) -> bytes:
    # restart interval 1 (no key prefix sharing) keeps the writer trivial
    body = bytearray()
    restarts = []
    for k, v in items:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(k)) + _put_varint(len(v)) + k + v
    for r in restarts:
        body += struct.pack('<I', r)
    body += struct.pack('<I', len(restarts))
    return bytes(body)


def _block_with_trailer(block: bytes) -> bytes:
    crc = _mask_crc(crc32c(block + b'\x00'))
    return block + b'\x00' + struct.pack('<I', crc)


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write ``tensors`` as ``prefix.index`` + ``prefix.data-00000-of-00001``."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors)
    offset = 0
    entries: List[Tuple[bytes, bytes]] = []
    # BundleHeaderProto{num_shards=1, endianness=LITTLE(0), version{producer=1}}
    header = _proto_field(1, 0, 1) + _proto_field(3, 2, _proto_field(1, 0, 1))
    entries.append((b'', header))
    # both files are written under temporary names and renamed into place, so that a reader (another rank waiting to
    # load the checkpoint, or a crash mid-write) never sees a truncated bundle
    tmp_data, tmp_index = prefix + '.data-00000-of-00001.tmp%d' % os.getpid(), prefix + '.index.tmp%d' % os.getpid()
    with open(tmp_data, 'wb') as f:
        for n in names:
            arr = np.asarray(tensors[n], order='C')
            entries.append((n.encode(), _entry_proto(arr, offset)))
            raw = arr.tobytes()
            f.write(raw)
            offset += len(raw)
    out = bytearray()
    data_block = _build_block(entries)
    data_handle = _put_varint(0) + _put_varint(len(data_block))
    out += _block_with_trailer(data_block)
    meta_block = _build_block([])
    meta_off = len(out)
    out += _block_with_trailer(meta_block)
    index_block = _build_block([(entries[-1][0] + b'\x00', data_handle)])
    index_off = len(out)
    out += _block_with_trailer(index_block)
    footer = (_put_varint(meta_off) + _put_varint(len(meta_block)) +
              _put_varint(index_off) + _put_varint(len(index_block)))
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', _MAGIC)
    out += footer
    with open(tmp_index, 'wb') as f:
        f.write(bytes(out))
    os.replace(tmp_data, prefix + '.data-00000-of-00001')
    os.replace(tmp_index, prefix + '.index')
