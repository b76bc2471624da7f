"""TensorFlow checkpoint (V2 "tensor bundle") reader / writer without TensorFlow.

The reference saves and restores with `tf.train.Saver` (reference trainer.py:180-213, 365-366; tester.py:260-309):
`model_dir/model.ckpt-<step>.{index,data-00000-of-00001}` plus the `checkpoint` state file, variables keyed by their
graph names (`Encoder/G_encoder/Conv_3/weights`, `ID_AE/G/...`, `Discriminator.2.Filters`, `PoseAE/...`, Adam slots
`<var>/Adam`, `<var>/Adam_1`, ...), restored partially by variable SCOPE.  The engines of this package keep their
parameters under exactly those names and layouts (HWIO / [in,out]), so importing a published DPIG checkpoint or
exporting one for the reference is a byte-level format question only.  TensorFlow is an un-vendored dependency of the
reference (README.md:17) and is not installable here, so the format is restated from its public definition:

  <prefix>.index    an SSTable (leveldb table format, tensorflow/core/lib/io/table*): prefix-compressed key/value
                    blocks (restart interval 16) each followed by a 1-byte compression tag + masked CRC32C, one index
                    block, an (empty) metaindex block and a 48-byte footer with magic 0xdb4775248b80fb57.
                    key ""   -> BundleHeaderProto {num_shards=1, endianness=2, version=3 {producer=1}}
                    key name -> BundleEntryProto  {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32)}
  <prefix>.data-SSSSS-of-NNNNN   raw little-endian tensor bytes at the recorded offsets (tensor_bundle.proto).

Parity note: no TensorFlow-written file is available in this environment (the reference ships none and there is no
network), so the writer is pinned only by this module's own reader, by the hand-assembled byte fixture in
tests/test_tf_checkpoint.py and by the format constants above -- "parity unpinned" in the sense of DESIGN.md §2.
"""
import os
import struct
from collections import OrderedDict

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8
_RESTART_INTERVAL = 16
_BLOCK_SIZE = 262144

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_ENUM = {np.dtype(v): k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------------ CRC32C
def _make_table():
    poly = 0x82F63B78
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab.append(c)
    return tab


_TABLE = _make_table()
_native_crc = None


def _load_native():
    """dpig_crc32c of libdpig.so (slicing-by-8, host code) for the bulk tensor bytes; the pure-Python loop below is
    used for the few-KB index blocks and whenever the library has not been built."""
    global _native_crc
    if _native_crc is None:
        try:
            import ctypes as C

            from . import _lib
            fn = _lib.load().dpig_crc32c
            fn.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
            fn.restype = C.c_uint32
            _native_crc = fn
        except Exception:
            _native_crc = False
    return _native_crc


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli, reflected 0x82F63B78) of a bytes-like object / contiguous numpy array."""
    if isinstance(data, np.ndarray):
        buf = np.frombuffer(np.asarray(data).tobytes(), dtype=np.uint8) if data.ndim == 0 else \
            np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    else:
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
    if buf.size >= 4096 and _load_native():
        return int(_native_crc(crc, buf.ctypes.data, buf.size))
    c = crc ^ 0xFFFFFFFF
    tab = _TABLE
    for b in buf.tobytes():
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    """leveldb / TensorFlow CRC masking (crc32c.h Mask): rotate right by 15 and add a constant."""
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ varints / protobuf
def _put_varint(out, v):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _parse_message(buf):
    """Minimal protobuf wire parser: returns {field: [values]} with varints as ints, fixed32/64 as ints and
    length-delimited fields as bytes."""
    out, pos, n = {}, 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _field(out, field, wt, payload):
    _put_varint(out, (field << 3) | wt)
    if wt == 0:
        _put_varint(out, payload)
    elif wt == 2:
        _put_varint(out, len(payload))
        out.extend(payload)
    elif wt == 5:
        out.extend(struct.pack("<I", payload))


def _encode_entry(dtype_enum, shape, shard_id, offset, size, crc_masked):
    """BundleEntryProto (tensor_bundle.proto); proto3 omits zero-valued scalars like TensorFlow's serializer does."""
    shp = bytearray()
    for d in shape:
        dim = bytearray()
        if d:
            _field(dim, 1, 0, int(d))
        _field(shp, 2, 2, dim)
    e = bytearray()
    _field(e, 1, 0, dtype_enum)
    _field(e, 2, 2, shp)
    if shard_id:
        _field(e, 3, 0, shard_id)
    if offset:
        _field(e, 4, 0, offset)
    if size:
        _field(e, 5, 0, size)
    _field(e, 6, 5, crc_masked)
    return bytes(e)


def _encode_header(num_shards=1):
    """BundleHeaderProto: num_shards, endianness LITTLE (=0, omitted), version {producer: 1}."""
    ver = bytearray()
    _field(ver, 1, 0, 1)
    h = bytearray()
    _field(h, 1, 0, num_shards)
    _field(h, 3, 2, ver)
    return bytes(h)


# ------------------------------------------------------------------------------------------ snappy (read side only)
def _snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
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
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        for _ in range(ln):      # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy: length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------------------------ SSTable
def _read_block(data, offset, size, verify=True):
    body = data[offset:offset + size]
    ctype = data[offset + size]
    stored = struct.unpack_from("<I", data, offset + size + 1)[0]
    if verify and mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
        raise ValueError("SSTable block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        body = _snappy_uncompress(body)
    elif ctype != 0:
        raise ValueError("unknown block compression %d" % ctype)
    return body


def _block_entries(block):
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as fh:
        data = fh.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError("%s is not an SSTable (bad magic)" % path)
    footer = data[len(data) - 48:]
    _, pos = _get_varint(footer, 0)          # metaindex handle
    _, pos = _get_varint(footer, pos)
    ioff, pos = _get_varint(footer, pos)
    isize, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p2 = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, p2)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


class _BlockBuilder:
    def __init__(self, restart_interval=None):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""
        self.interval = restart_interval or _RESTART_INTERVAL

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        _put_varint(self.buf, shared)
        _put_varint(self.buf, len(key) - shared)
        _put_varint(self.buf, len(value))
        self.buf += key[shared:]
        self.buf += value
        self.last = key
        self.count += 1

    def finish(self):
        out = bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))
        return out

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return not self.buf


def _shortest_separator(a, b):
    """leveldb BytewiseComparator::FindShortestSeparator: a short key k with a <= k < b."""
    m = min(len(a), len(b))
    i = 0
    while i < m and a[i] == b[i]:
        i += 1
    if i < m and a[i] < 0xFF and a[i] + 1 < b[i]:
        return a[:i] + bytes([a[i] + 1])
    return a


def _short_successor(a):
    for i, c in enumerate(a):
        if c != 0xFF:
            return a[:i] + bytes([c + 1])
    return a


def write_table(path, items):
    """items: (key bytes, value bytes) sorted by key.  Uncompressed blocks, like TensorFlow's BundleWriter."""
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)    # TensorFlow's table builder: index_block(index_block_restart_interval = 1)
    pending = None   # (last key of the finished block, handle)

    def emit(block_bytes):
        off = len(out)
        out.extend(block_bytes)
        out.append(0)
        out.extend(struct.pack("<I", mask_crc(crc32c(block_bytes + b"\x00"))))
        h = bytearray()
        _put_varint(h, off)
        _put_varint(h, len(block_bytes))
        return bytes(h)

    blk = _BlockBuilder()
    for key, value in items:
        if pending is not None:
            index.add(_shortest_separator(pending[0], key), pending[1])
            pending = None
        blk.add(key, value)
        if blk.size() >= _BLOCK_SIZE:
            pending = (blk.last, emit(blk.finish()))
            blk = _BlockBuilder()
    if not blk.empty():
        pending = (blk.last, emit(blk.finish()))
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    meta_handle = emit(_BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = bytearray(meta_handle + index_handle)
    footer.extend(b"\x00" * (40 - len(footer)))
    footer.extend(struct.pack("<Q", _MAGIC))
    out.extend(footer)
    with open(path, "wb") as fh:
        fh.write(out)


# ------------------------------------------------------------------------------------------ bundle API
def _data_path(prefix, shard, num_shards):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


class CheckpointReader:
    """Same surface as `tf.train.NewCheckpointReader(prefix)`: has_tensor / get_tensor /
    get_variable_to_shape_map / get_variable_to_dtype_map."""

    def __init__(self, prefix, verify=True):
        self.prefix, self.verify = prefix, verify
        self.entries = OrderedDict()
        self.num_shards = 1
        for key, value in read_table(prefix + ".index", verify):
            msg = _parse_message(value)
            if key == b"":
                self.num_shards = msg.get(1, [1])[0]
                if msg.get(2, [0])[0] != 0:
                    raise ValueError("big-endian checkpoints are not supported")
                continue
            shape = []
            for shp in msg.get(2, []):
                for dim in _parse_message(shp).get(2, []):
                    shape.append(_parse_message(dim).get(1, [0])[0])
            if 7 in msg:
                raise ValueError("partitioned variable %r (tensor slices) is not supported" % key.decode())
            self.entries[key.decode()] = dict(dtype=msg.get(1, [0])[0], shape=tuple(shape), shard=msg.get(3, [0])[0],
                                              offset=msg.get(4, [0])[0], size=msg.get(5, [0])[0], crc=msg.get(6, [0])[0])

    def has_tensor(self, name):
        return name in self.entries

    def get_variable_to_shape_map(self):
        return {k: list(e["shape"]) for k, e in self.entries.items()}

    def get_variable_to_dtype_map(self):
        return {k: np.dtype(_DTYPES[e["dtype"]]) for k, e in self.entries.items()}

    def get_tensor(self, name):
        e = self.entries[name]
        if e["dtype"] not in _DTYPES:
            raise ValueError("tensor %r has unsupported dtype enum %d" % (name, e["dtype"]))
        dt = np.dtype(_DTYPES[e["dtype"]])
        with open(_data_path(self.prefix, e["shard"], self.num_shards), "rb") as fh:
            fh.seek(e["offset"])
            raw = fh.read(e["size"])
        if len(raw) != e["size"] or e["size"] != int(np.prod(e["shape"], dtype=np.int64)) * dt.itemsize:
            raise ValueError("tensor %r: size mismatch" % name)
        arr = np.frombuffer(raw, dtype=dt).reshape(e["shape"])
        if self.verify and mask_crc(crc32c(arr)) != e["crc"]:
            raise ValueError("tensor %r: checksum mismatch" % name)
        return arr.copy()


def load_checkpoint(prefix, scopes=None, verify=True):
    """dict name -> array of every variable (or of those whose name starts with one of `scopes`: the reference's
    partial restores `tf.train.Saver(var_list=<variables of a scope>)`, trainer.py:180-187, tester.py:260-277)."""
    r = CheckpointReader(prefix, verify)
    out = OrderedDict()
    for name in r.entries:
        if scopes is None or any(name.startswith(s) for s in scopes):
            out[name] = r.get_tensor(name)
    return out


def save_checkpoint(prefix, tensors, update_state=True):
    """Writes <prefix>.index / <prefix>.data-00000-of-00001 (and the `checkpoint` state file next to them, like
    tf.train.Saver.save).  tensors: dict name -> array; names and layouts are taken as they are."""
    names = sorted(tensors, key=lambda s: s.encode())
    items = [(b"", _encode_header(1))]
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    offset = 0
    with open(_data_path(prefix, 0, 1), "wb") as fh:
        for name in names:
            arr = np.asarray(tensors[name])
            if arr.ndim and not arr.flags.c_contiguous:
                arr = np.ascontiguousarray(arr)
            if arr.dtype not in _DTYPE_ENUM:
                raise ValueError("tensor %r: dtype %s has no TensorFlow counterpart here" % (name, arr.dtype))
            fh.write(arr.tobytes())
            items.append((name.encode(), _encode_entry(_DTYPE_ENUM[arr.dtype], arr.shape, 0, offset, arr.nbytes,
                                                       mask_crc(crc32c(arr)))))
            offset += arr.nbytes
    write_table(prefix + ".index", items)
    if update_state:
        d = os.path.dirname(os.path.abspath(prefix))
        state = os.path.join(d, "checkpoint")
        base = os.path.basename(prefix)
        prev = []
        if os.path.exists(state):
            for ln in open(state):
                if ln.startswith("all_model_checkpoint_paths:"):
                    prev.append(ln.split(":", 1)[1].strip().strip('"'))
        prev = [p for p in prev if p != base] + [base]
        with open(state, "w") as fh:
            fh.write('model_checkpoint_path: "%s"\n' % base)
            for p in prev:
                fh.write('all_model_checkpoint_paths: "%s"\n' % p)
    return prefix


def latest_checkpoint(model_dir):
    """tf.train.latest_checkpoint: the prefix named by model_dir/checkpoint, or None."""
    state = os.path.join(model_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    for ln in open(state):
        if ln.startswith("model_checkpoint_path:"):
            p = ln.split(":", 1)[1].strip().strip('"')
            return p if os.path.isabs(p) else os.path.join(model_dir, p)
    return None


def is_checkpoint(path):
    return bool(path) and os.path.exists(path + ".index")


def load_any(path, scopes=None):
    """Parameters from a TensorFlow checkpoint prefix, a directory holding a `checkpoint` state file, or an .npz."""
    if path and os.path.isdir(path):
        path = latest_checkpoint(path) or path
    if is_checkpoint(path):
        return load_checkpoint(path, scopes)
    with np.load(path) as z:
        return OrderedDict((k, z[k]) for k in z.files if scopes is None or any(k.startswith(s) for s in scopes))
