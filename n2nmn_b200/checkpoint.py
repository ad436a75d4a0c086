"""TensorFlow checkpoint import / export for the module-network variables.

The reference saves and restores with ``tf.train.Saver`` (exp_clevr/train_clevr_gt_layout.py:
158-159,219-223; exp_clevr/eval_clevr.py:90-91). TensorFlow >= 0.12 (the reference pins 1.0.0)
writes the "V2" tensor-bundle format:

    <prefix>.index                 an SSTable (LevelDB table format): key "" -> BundleHeaderProto,
                                   key <variable name> -> BundleEntryProto
                                   {dtype, shape, shard_id, offset, size, crc32c}
    <prefix>.data-00000-of-00001   the raw little-endian tensor bytes, back to back

TensorFlow is not installable here, so the two file formats are restated from their public
specification (tensorflow/core/lib/io/{format,block,table}.cc, tensorflow/core/util/tensor_bundle,
tensor_bundle.proto): blocks of prefix-compressed entries with restart arrays, a 5-byte block
trailer (compression type + masked CRC-32C), a 48-byte footer with the metaindex / index block
handles and the magic 0xdb4775248b80fb57. ``read_checkpoint`` accepts uncompressed and
Snappy-compressed blocks; ``write_checkpoint`` emits uncompressed blocks with valid checksums, the
layout ``tf.train.Saver.restore`` / ``tf.train.load_checkpoint`` expect.

Variable names: the reference's module variables live under
``neural_module_network/layout_execution/module_variables/<Scope>/<layer>/{weights,biases}``
(models_clevr/nmn3_model.py:49-52, nmn3_modules.py:17-18); ``import_module_weights`` strips that
prefix and ignores optimizer slots (``.../Adam``, ``.../Adam_1``, ``beta1_power`` ...) and the
seq2seq variables, returning exactly what ``Modules.set_weights`` takes.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

MODULE_SCOPE = 'neural_module_network/layout_execution/module_variables/'
_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8,
           9: np.int64, 10: np.bool_, 17: np.uint16, 19: np.float16}
_DTYPE_ENUM = {np.dtype(v): k for k, v in _DTYPES.items()}


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli) through the library's host helper (n2nmn_crc32c)."""
    from . import _lib
    buf = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    return int(_lib.lib().n2nmn_crc32c(C.c_char_p(bytes(buf)), len(buf), crc))


def _mask(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + _MASK_DELTA) & 0xffffffff


def _unmask(m):
    rot = (m - _MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ------------------------------------------------------------------------------ varints / protobuf
def _get_varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7f) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """Generic protobuf wire decode: {field: [values]}; length-delimited fields stay bytes."""
    pos, out = 0, {}
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        out.setdefault(field, []).append(v)
    return out


def _field(field, wt, payload):
    return _put_varint((field << 3) | wt) + payload


def _parse_entry(buf):
    """BundleEntryProto: dtype=1, shape=2 (TensorShapeProto: dim=2 {size=1}), shard_id=3, offset=4,
    size=5, crc32c=6 (fixed32)."""
    p = _parse_proto(buf)
    shape = []
    for sp in p.get(2, []):
        for dim in _parse_proto(sp).get(2, []):
            shape.append(_parse_proto(dim).get(1, [0])[0])
    return dict(dtype=p.get(1, [0])[0], shape=tuple(shape), shard_id=p.get(3, [0])[0],
                offset=p.get(4, [0])[0], size=p.get(5, [0])[0], crc32c=p.get(6, [None])[0])


def _encode_entry(dtype_enum, shape, offset, size, crc):
    dims = b''.join(_field(2, 2, _put_varint(len(d)) + d)
                    for d in (_field(1, 0, _put_varint(int(s))) for s in shape))
    out = _field(1, 0, _put_varint(dtype_enum))
    out += _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        out += _field(4, 0, _put_varint(offset))
    out += _field(5, 0, _put_varint(size))
    out += _field(6, 5, struct.pack('<I', crc))
    return out


# ------------------------------------------------------------------------------ snappy (decode only)
def _snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], 'little')
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        for _ in range(ln):                             # may overlap: byte by byte
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('corrupt snappy block')
    return bytes(out)


# ------------------------------------------------------------------------------ table (SSTable)
def _read_block(data, offset, size, verify=True):
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    stored = struct.unpack_from('<I', data, offset + size + 1)[0]
    if verify and _unmask(stored) != crc32c(data[offset:offset + size + 1]):
        raise ValueError('checkpoint index: block checksum mismatch at offset %d' % offset)
    if ctype == 1:
        raw = _snappy_decompress(raw)
    elif ctype != 0:
        raise ValueError('checkpoint index: unknown block compression %d' % ctype)
    return raw


def _block_entries(block):
    num_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _table_items(data, verify=True):
    if len(data) < 48 or struct.unpack_from('<Q', data, len(data) - 8)[0] != _MAGIC:
        raise ValueError('not a TensorFlow checkpoint index (bad table magic)')
    footer = data[len(data) - 48:]
    pos = 0
    _, pos = _get_varint(footer, pos)        # metaindex handle
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p2 = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, p2)
        for kv in _block_entries(_read_block(data, boff, bsize, verify)):
            yield kv


def _build_block(items, restart_interval):
    out, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(last), len(k))
            while shared < m and last[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v))
        out += k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _emit_block(f, block):
    off = f.tell()
    f.write(block)
    f.write(b'\x00' + struct.pack('<I', _mask(crc32c(block + b'\x00'))))
    return _put_varint(off) + _put_varint(len(block))


# ------------------------------------------------------------------------------ public API
def read_checkpoint(prefix, verify=True):
    """{variable name: numpy array} of every tensor in a V2 checkpoint `<prefix>.index` +
    `<prefix>.data-*`."""
    with open(prefix + '.index', 'rb') as f:
        index = f.read()
    entries, num_shards = {}, 1
    for key, val in _table_items(index, verify):
        if key == b'':
            num_shards = _parse_proto(val).get(1, [1])[0]
        else:
            entries[key.decode()] = _parse_entry(val)
    shards = {}
    out = {}
    for name, e in entries.items():
        if e['dtype'] not in _DTYPES:
            continue                                    # strings etc.: not model variables
        sid = e['shard_id']
        if sid not in shards:
            with open('%s.data-%05d-of-%05d' % (prefix, sid, num_shards), 'rb') as f:
                shards[sid] = f.read()
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if verify and e['crc32c'] is not None and _unmask(e['crc32c']) != crc32c(raw):
            raise ValueError('checkpoint data: checksum mismatch for %s' % name)
        out[name] = np.frombuffer(raw, dtype=_DTYPES[e['dtype']]).reshape(e['shape']).copy()
    return out


def write_checkpoint(prefix, tensors):
    """Writes {name: array} as a single-shard V2 checkpoint that tf.train.Saver can restore."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(tensors, key=lambda s: s.encode())
    items = [(b'', _field(1, 0, _put_varint(1)) +                   # num_shards = 1
              _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1))))]   # version.producer = 1
    offset = 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for name in names:
            a = np.asarray(tensors[name])
            a = a if a.flags.c_contiguous else np.array(a, order='C')
            if a.dtype not in _DTYPE_ENUM:
                raise TypeError('unsupported dtype %s for %s' % (a.dtype, name))
            raw = a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes()
            f.write(raw)
            items.append((name.encode(), _encode_entry(_DTYPE_ENUM[a.dtype], a.shape, offset,
                                                       len(raw), _mask(crc32c(raw)))))
            offset += len(raw)
    with open(prefix + '.index', 'wb') as f:
        handles, block_items = [], []
        for i, kv in enumerate(items):                  # data blocks of <= 64 entries
            block_items.append(kv)
            if len(block_items) == 64 or i == len(items) - 1:
                handles.append((block_items[-1][0], _emit_block(f, _build_block(block_items, 16))))
                block_items = []
        meta = _emit_block(f, _build_block([], 1))
        index = _emit_block(f, _build_block(handles, 1))
        footer = meta + index
        f.write(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', _MAGIC))


def import_module_weights(prefix, family='clevr', H=10, W=15, D=512, num_choices=28,
                          text_dim=300):
    """Module-network weights of a reference snapshot as {relative name: float32 array}, shapes
    checked against the family's variable table (weights.variable_shapes). Returns
    (weights, ignored_names)."""
    from .weights import variable_shapes
    want = variable_shapes(family, H, W, D, num_choices, text_dim)
    got, ignored = {}, []
    for name, arr in read_checkpoint(prefix).items():
        i = name.find(MODULE_SCOPE)
        rel = name[i + len(MODULE_SCOPE):] if i >= 0 else None
        if rel in want:
            if tuple(arr.shape) != tuple(want[rel]):
                raise ValueError('%s: checkpoint shape %s, model expects %s'
                                 % (name, arr.shape, tuple(want[rel])))
            got[rel] = arr.astype(np.float32)
        else:
            ignored.append(name)
    missing = sorted(set(want) - set(got))
    if missing:
        raise KeyError('checkpoint %s lacks module variables: %s' % (prefix, missing))
    return got, sorted(ignored)


def export_module_weights(prefix, weights, extra=None):
    """Saves {relative name: array} under the reference's TF variable names (+ `extra` tensors
    under their own names), so the snapshot restores into the reference graph."""
    tensors = {MODULE_SCOPE + k: np.asarray(v, np.float32) for k, v in weights.items()}
    tensors.update(extra or {})
    write_checkpoint(prefix, tensors)
