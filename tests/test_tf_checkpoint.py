"""CPU tests of the TensorFlow-free checkpoint reader / writer (tf_checkpoint.py; reference tf.train.Saver use at
trainer.py:180-213, 365-366 and tester.py:260-309).  No TensorFlow-written file exists in this environment, so the
format is pinned by (a) known answers of its primitives (CRC-32C check value, leveldb CRC mask, varints), (b) a tiny
index file assembled BYTE BY BYTE in this test from the published format description, which the writer must reproduce
exactly and the reader must parse, and (c) round trips at realistic sizes / names."""
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import dpig_b200  # noqa: E402,F401
from dpig_b200 import tf_checkpoint as ck  # noqa: E402


def test_crc32c_known_answers():
    assert ck.crc32c(b"123456789") == 0xE3069283                    # the CRC-32C check value
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(bytes(32)) == 0x8A9136AA                       # RFC 3720 B.4: 32 bytes of zeros
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43              # RFC 3720 B.4: 32 bytes of ones
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E                # RFC 3720 B.4: incrementing bytes
    # native (libdpig.so, slicing-by-8) path == pure-Python path, incl. unaligned starts and chaining
    rng = np.random.default_rng(0)
    buf = rng.integers(0, 256, size=70001, dtype=np.uint8)
    py = 0xFFFFFFFF
    for b in buf.tobytes():
        py = ck._TABLE[(py ^ b) & 0xFF] ^ (py >> 8)
    py ^= 0xFFFFFFFF
    assert ck.crc32c(buf) == py
    assert ck.crc32c(buf[3:]) == ck.crc32c(buf[3:].tobytes()[:100] + buf[103:].tobytes())
    a, b = buf[:50000], buf[50000:]
    assert ck.crc32c(b, ck.crc32c(a)) == py
    # leveldb mask: ((crc >> 15) | (crc << 17)) + 0xa282ead8
    assert ck.mask_crc(0) == 0xA282EAD8 and ck.mask_crc(0xE3069283) == (((0xE3069283 >> 15) | (0xE3069283 << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _block(entries):
    """leveldb block: (shared, non_shared, value_len, key delta, value)* restarts[] num_restarts; one restart at 0."""
    body = b"".join(bytes([sh, ns, len(v)]) + kd + v for sh, ns, kd, v in entries)
    return body + struct.pack("<II", 0, 1)


def _with_trailer(block):
    return block + b"\x00" + struct.pack("<I", ck.mask_crc(ck.crc32c(block + b"\x00")))


def test_writer_reproduces_hand_assembled_bundle(tmp_path):
    """One float32 [2,3] variable 'w': every byte of the expected .index is written out here by hand."""
    w = np.arange(6, dtype=np.float32).reshape(2, 3)
    prefix = str(tmp_path / "model.ckpt-7")
    ck.save_checkpoint(prefix, {"w": w})
    assert open(prefix + ".data-00000-of-00001", "rb").read() == w.tobytes()
    # BundleHeaderProto: field1 varint num_shards=1 -> 08 01 ; field3 len-delimited VersionDef{field1 producer=1} -> 1a 02 08 01
    header = bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    # BundleEntryProto: dtype DT_FLOAT=1 -> 08 01 ; shape {dim{size:2} dim{size:3}} -> 12 08 (12 02 08 02)(12 02 08 03);
    #                   offset 0 omitted ; size=24 -> 28 18 ; crc32c fixed32 -> 35 xx xx xx xx
    entry = bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x02, 0x12, 0x02, 0x08, 0x03, 0x28, 0x18, 0x35]) + \
        struct.pack("<I", ck.mask_crc(ck.crc32c(w.tobytes())))
    data_block = _block([(0, 0, b"", header), (0, 1, b"w", entry)])
    meta_block = _block([])
    # index block: one entry, key = short successor of the last key 'w' = 'x', value = BlockHandle(offset 0, size)
    d_len, m_off = len(data_block), len(data_block) + 5
    assert d_len < 128 and m_off < 128           # single-byte varints below
    index_block = _block([(0, 1, b"x", bytes([0, d_len]))])
    i_off = m_off + len(meta_block) + 5
    footer = bytes([m_off, len(meta_block), i_off, len(index_block)])
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    expected = _with_trailer(data_block) + _with_trailer(meta_block) + _with_trailer(index_block) + footer
    assert open(prefix + ".index", "rb").read() == expected
    r = ck.CheckpointReader(prefix)
    assert r.get_variable_to_shape_map() == {"w": [2, 3]} and r.get_variable_to_dtype_map()["w"] == np.float32
    assert np.array_equal(r.get_tensor("w"), w)
    assert open(str(tmp_path / "checkpoint")).read().splitlines()[0] == 'model_checkpoint_path: "model.ckpt-7"'
    assert ck.latest_checkpoint(str(tmp_path)) == prefix


def test_round_trip_with_reference_variable_names(tmp_path):
    """All Stage-I variables of a reduced graph + Adam slots + scalars, several index blocks' worth of keys, scope-partial
    restore like tf.train.Saver(var_list=<scope variables>) (trainer.py:180-187)."""
    from dpig_b200 import engine
    cfg = engine.NetConfig(img_h=32, img_w=16, hidden=8, roi_size=12, d_dim=8)
    params = engine.init_params(cfg, seed=3)
    tensors = dict(params)
    for k, v in params.items():
        tensors[k + "/Adam"] = np.zeros_like(v)
        tensors[k + "/Adam_1"] = np.ones_like(v)
    tensors["step"] = np.array(1234, dtype=np.int32)
    tensors["beta1_power"] = np.array(0.5, dtype=np.float32)
    tensors["g_lr"] = np.array(2e-5, dtype=np.float64)
    tensors["flags"] = np.array([True, False, True])
    tensors["ids"] = np.arange(5, dtype=np.int64)
    prefix = str(tmp_path / "sub" / "model.ckpt-99")
    ck.save_checkpoint(prefix, tensors)
    r = ck.CheckpointReader(prefix)
    assert list(r.entries) == sorted(tensors, key=lambda s: s.encode())          # SSTable key order
    for k, v in tensors.items():
        got = r.get_tensor(k)
        assert got.dtype == np.asarray(v).dtype and got.shape == np.asarray(v).shape and np.array_equal(got, v), k
    enc = ck.load_checkpoint(prefix, scopes=["Encoder/", "ID_AE/"])
    assert enc and all(k.startswith(("Encoder/", "ID_AE/")) for k in enc)
    assert set(k for k in enc if not k.endswith(("/Adam", "/Adam_1"))) == set(k for k in params if not k.startswith("Discriminator"))
    # a second save in the same directory extends the state file (Saver keeps the list of recent checkpoints)
    ck.save_checkpoint(str(tmp_path / "sub" / "model.ckpt-199"), {"step": np.array(5, np.int32)})
    lines = open(str(tmp_path / "sub" / "checkpoint")).read().splitlines()
    assert lines == ['model_checkpoint_path: "model.ckpt-199"', 'all_model_checkpoint_paths: "model.ckpt-99"',
                     'all_model_checkpoint_paths: "model.ckpt-199"']
    assert ck.load_any(str(tmp_path / "sub"))["step"] == 5


def test_many_keys_span_several_blocks_and_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {"scope_%04d/some/long/variable/name/weights" % i: rng.normal(size=(3, 5)).astype(np.float32)
               for i in range(6000)}                                                    # > 256 KB of index entries
    prefix = str(tmp_path / "big")
    ck.save_checkpoint(prefix, tensors, update_state=False)
    assert len(ck.read_table(prefix + ".index")) == 6001
    got = ck.load_checkpoint(prefix)
    assert all(np.array_equal(got[k], v) for k, v in tensors.items())
    # restart points: prefix compression restarts every 16 keys, keys still reconstruct
    assert list(got) == sorted(tensors)
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[100] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        ck.load_checkpoint(prefix)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[50] ^= 1
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="checksum"):
        ck.read_table(prefix + ".index")


def test_snappy_blocks_are_readable():
    # literal "abcd" + copy(offset 4, len 8) + literal "xyz"  -> "abcdabcdabcdxyz"
    comp = bytes([15, (4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([(3 - 1) << 2]) + b"xyz"
    assert ck._snappy_uncompress(comp) == b"abcdabcdabcdxyz"
