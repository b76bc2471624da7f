"""CPU tests of the TFRecord / tf.Example reader that stands where datasets/market1501.py:50-162 + trainer.py:537-564
(_load_batch_pair_pose) stand in the reference.  No TensorFlow-written file exists here, so the format is pinned by

  * a record and an Example assembled BY HAND from the published wire formats (record_writer.cc framing, protobuf
    varints / length-delimited fields, example.proto field numbers) -- independent of this module's writer,
  * the RFC 3720 CRC-32C known answers (tests/test_tf_checkpoint.py covers those),
  * writer -> reader round trips and corruption detection.
"""
import importlib
import io
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("disentangled-person-image-generation_b200")
ds = importlib.import_module("disentangled-person-image-generation_b200.datasets")
tfc = importlib.import_module("disentangled-person-image-generation_b200.tf_checkpoint")
synth = importlib.import_module("disentangled-person-image-generation_b200.synth")


def _hand_example():
    """Example{features{ feature{"a": int64_list [3, 300, -1] (packed)}, feature{"b": float_list [1.5] (unpacked)},
    feature{"s": bytes_list ["hi", ""]} }} byte by byte."""
    neg1 = b"\xff" * 9 + b"\x01"                                   # varint of 2^64-1
    ints = b"\x03" + b"\xac\x02" + neg1                            # 3, 300, -1
    int64_list = b"\x0a" + bytes([len(ints)]) + ints               # Int64List.value = 1, packed
    feat_a = b"\x1a" + bytes([len(int64_list)]) + int64_list       # Feature.int64_list = 3
    float_list = b"\x0d" + struct.pack("<f", 1.5)                  # FloatList.value = 1, wire type 5 (unpacked)
    feat_b = b"\x12" + bytes([len(float_list)]) + float_list       # Feature.float_list = 2
    bytes_list = b"\x0a\x02hi" + b"\x0a\x00"                       # BytesList.value = 1 twice
    feat_s = b"\x0a" + bytes([len(bytes_list)]) + bytes_list       # Feature.bytes_list = 1

    def entry(key, feat):
        e = b"\x0a" + bytes([len(key)]) + key + b"\x12" + bytes([len(feat)]) + feat
        return b"\x0a" + bytes([len(e)]) + e                       # Features.feature = 1 (map entry)

    feats = entry(b"a", feat_a) + entry(b"b", feat_b) + entry(b"s", feat_s)
    return b"\x0a" + bytes([len(feats)]) + feats                   # Example.features = 1


def test_parse_hand_assembled_example():
    ex = ds.parse_example(_hand_example())
    assert ex["a"].dtype == np.int64 and ex["a"].tolist() == [3, 300, -1]
    assert ex["b"].dtype == np.float32 and ex["b"].tolist() == [1.5]
    assert ex["s"] == [b"hi", b""]


def test_record_framing_hand_assembled(tmp_path):
    # TFRecord framing: u64 length | masked crc32c(length bytes) | payload | masked crc32c(payload)
    payload = b"123456789"
    assert tfc.crc32c(payload) == 0xE3069283                      # RFC 3720 B.4 check value
    head = struct.pack("<Q", 9)
    rot = lambda c: ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    blob = head + struct.pack("<I", rot(tfc.crc32c(head))) + payload + struct.pack("<I", rot(0xE3069283))
    f = tmp_path / "one.tfrecord"
    f.write_bytes(blob + blob)
    assert list(ds.read_records(str(f))) == [payload, payload]
    g = tmp_path / "w.tfrecord"
    ds.write_records(str(g), [payload, payload])
    assert g.read_bytes() == blob + blob                           # the writer emits exactly the hand-built bytes


def test_corruption_and_truncation_detected(tmp_path):
    f = tmp_path / "c.tfrecord"
    ds.write_records(str(f), [b"abcdefgh" * 8, b"xyz"])
    raw = bytearray(f.read_bytes())
    bad = bytearray(raw)
    bad[20] ^= 0x01                                                # payload bit flip
    f.write_bytes(bad)
    with pytest.raises(ValueError, match="payload"):
        list(ds.read_records(str(f)))
    assert len(list(ds.read_records(str(f), verify=False))) == 2   # tf.python_io's reader has the same opt-out
    bad = bytearray(raw)
    bad[3] ^= 0x01                                                 # length bit flip
    f.write_bytes(bad)
    with pytest.raises(ValueError, match="length"):
        list(ds.read_records(str(f)))
    f.write_bytes(raw[:-2])
    with pytest.raises(ValueError, match="truncated"):
        list(ds.read_records(str(f)))
    f.write_bytes(b"")
    assert list(ds.read_records(str(f))) == []                     # empty file = no records


def test_varints_vectorised_matches_scalar():
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.integers(0, 2, 500), rng.integers(-2 ** 62, 2 ** 62, 300), [0, 127, 128, 16383, 16384, -1]])
    buf = bytearray()
    for v in vals.tolist():
        tfc._put_varint(buf, int(v))
    assert ds._varints(bytes(buf)).tolist() == vals.tolist()
    assert ds._varints(b"").size == 0
    small = bytes(rng.integers(0, 2, 8192).astype(np.uint8))      # the {0,1} mask fast path
    assert ds._varints(small).tolist() == list(small)


def test_example_round_trip_types():
    f = {"f": np.array([0.25, -3.0, 1e-8], np.float32), "i": np.array([-5, 0, 2 ** 40], np.int64), "raw": b"\x00\xff\x10",
         "names": [b"a", b"bc"], "empty_f": np.zeros(0, np.float32)}
    ex = ds.parse_example(ds.encode_example(f))
    assert np.array_equal(ex["f"], f["f"]) and np.array_equal(ex["i"], f["i"])
    assert ex["raw"] == [f["raw"]] and ex["names"] == f["names"] and ex["empty_f"].size == 0


def _pairs(n, h, w, seed):
    b = synth.make_batch(n, h, w, seed=seed)
    t = synth.make_batch(n, h, w, seed=seed + 1000)
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        # smooth images so that JPEG at quality 100 stays within a few grey levels
        base = rng.integers(40, 215, size=(h // 8, w // 8, 3)).astype(np.uint8)
        img = np.kron(base, np.ones((8, 8, 1), np.uint8))
        out.append(dict(x=img, x_target=img[:, ::-1].copy(), pose_rcv=b["pose_rcv"][i], pose_rcv_target=t["pose_rcv"][i],
                        mask=b["mask"][i, :, :, 0], mask_target=t["mask"][i, :, :, 0], part_bbox=b["part_bbox"][i],
                        part_bbox_target=t["part_bbox"][i], part_vis=b["part_vis"][i], part_vis_target=t["part_vis"][i],
                        label=i % 2, name="p%d" % i))
    return out


@pytest.mark.parametrize("data_name", ["Market1501", "DeepFashion"])
def test_loader_batches_match_reference_shapes(tmp_path, data_name):
    h, w = ds.DATASETS[data_name]
    n = 6 if data_name == "Market1501" else 3
    pairs = _pairs(n, h, w, seed=5)
    ds.write_pair_records(str(tmp_path / ("%s_train_00000-of-00002.tfrecord" % data_name)), pairs[: n // 2], quality=100, subsampling=0)
    ds.write_pair_records(str(tmp_path / ("%s_train_00001-of-00002.tfrecord" % data_name)), pairs[n // 2:], quality=100, subsampling=0)
    ld = ds.get_split("train", str(tmp_path), data_name=data_name, batch_size=n, shuffle=False, repeat=False)
    b = ld.next_batch()
    # shapes / dtypes / ranges of _load_batch_pair_pose (trainer.py:544-558)
    assert b["x"].shape == (n, h, w, 3) and b["x"].dtype == np.float32 and abs(b["x"]).max() <= 1.0
    assert b["x_target"].shape == (n, h, w, 3)
    assert b["pose_rcv"].shape == (n, 18, 3) and b["mask"].shape == (n, h, w, 1) and b["mask_r4"].shape == (n, h, w, 1)
    assert b["part_bbox"].shape == (n, 37, 4) and b["part_bbox"].dtype == np.int64 and b["part_vis"].shape == (n, 37)
    for i, p in enumerate(pairs):                                  # file order, record order preserved without shuffle
        assert np.array_equal(b["pose_rcv"][i], p["pose_rcv"]) and np.array_equal(b["pose_rcv_target"][i], p["pose_rcv_target"])
        assert np.array_equal(b["mask"][i, :, :, 0], p["mask"]) and np.array_equal(b["part_bbox"][i], p["part_bbox"])
        assert np.array_equal(b["part_vis_target"][i], p["part_vis_target"]) and b["label"][i] == i % 2
        assert np.abs(b["x"][i] * 127.5 + 127.5 - p["x"]).max() <= 6.0      # JPEG q100 of a blocky image
    with pytest.raises(StopIteration):
        ld.next_batch()


def test_loader_shuffle_repeat_and_errors(tmp_path):
    pairs = _pairs(5, 128, 64, seed=9)
    ds.write_pair_records(str(tmp_path / "Market1501_train_00000-of-00001.tfrecord"), pairs)
    ld = ds.get_split("train", str(tmp_path), batch_size=4, shuffle=True, seed=1)
    seen = set()
    for _ in range(12):                                            # repeat=True: an endless stream, like the TF queue
        seen.update(int(v) for v in ld.next_batch()["part_bbox"][:, 0, 2])
    ld.close()
    assert seen == {int(p["part_bbox"][0, 2]) for p in pairs}
    with pytest.raises(ValueError):
        ds.get_split("val", str(tmp_path))                         # market1501.py:68-69
    with pytest.raises(IOError):
        ds.get_split("test", str(tmp_path))                        # no files of that split
    bad = dict(pairs[0], x=np.zeros((64, 64, 3), np.uint8))
    ds.write_pair_records(str(tmp_path / "Market1501_test_00000-of-00001.tfrecord"), [bad])
    with pytest.raises(ValueError, match="expected 128x64x3"):     # producer-thread errors surface in the consumer
        ds.get_split("test", str(tmp_path), batch_size=1, shuffle=False).next_batch()


def test_trainer_make_loader_routes_to_tfrecords(tmp_path):
    cfgm = importlib.import_module("disentangled-person-image-generation_b200.config")
    tr = importlib.import_module("disentangled-person-image-generation_b200.trainer")
    d = tmp_path / "Market_train_data"
    d.mkdir()
    ds.write_pair_records(str(d / "Market1501_train_00000-of-00001.tfrecord"), _pairs(2, 128, 64, seed=2))
    cfg, _ = cfgm.get_config(["--dataset=Market_train_data", "--data_dir=%s" % tmp_path, "--synthetic_data=false",
                              "--batch_size=2"])
    ld = tr.make_loader(cfg, 2, 128, 64)
    assert isinstance(ld, ds.TFRecordPairLoader) and ld.next_batch()["x"].shape == (2, 128, 64, 3)
    ld.close()
    cfg, _ = cfgm.get_config(["--dataset=Market_train_data"])
    assert isinstance(tr.make_loader(cfg, 2, 128, 64), tr.SyntheticLoader)
