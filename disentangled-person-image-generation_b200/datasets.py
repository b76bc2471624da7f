"""Input pipeline of the reference without TensorFlow: TFRecord files of tf.Example protos in the Market-1501 /
DeepFashion pair schema -> host batches for `Stage1Engine.set_batch`.

    datasets/market1501.py:50-162, datasets/deepfashion.py     get_split(): file pattern, feature keys, shapes
    trainer.py:537-564  _load_batch_pair_pose()                which keys reach the graph, reshape / cast / normalise
    trainer.py:105-110, 553-555                                queue runners: a background producer, `tf.train.batch`

The reference's converters (datasets/convert_market.py, convert_DF.py) write one tf.Example per image PAIR with the keys
listed in FEATURE_KEYS below; `<data_name>_<split>_*.tfrecord` under the dataset directory.  This module restates

  * the TFRecord framing (tensorflow/core/lib/io/record_writer.cc): uint64 length, masked CRC-32C of the length,
    payload, masked CRC-32C of the payload (CRC via libdpig.so's dpig_crc32c / tf_checkpoint.crc32c);
  * the tf.Example wire format (example.proto / feature.proto): Example{features=1}, Features{map feature=1},
    Feature{bytes_list=1 | float_list=2 | int64_list=3}, lists packed or unpacked;
  * the batch assembly of _load_batch_pair_pose: x = (uint8 image - 127.5)/127.5 NHWC, pose_rcv [18,3], mask_r6 [H,W,1],
    part_bbox [37,4], part_vis [37] for both images of the pair (`*_target` = image 1).  The pose MAPS are rasterised
    on the GPU by dpig_pose_rasterize from pose_rcv, so they are not produced here.

JPEG decoding uses PIL (the reference: tf.image.decode_jpeg inside slim's tfexample_decoder.Image).
Writer functions exist for tests and for exporting synthetic data in the reference's format; no TensorFlow-written
record file is available in this environment, so the reader is pinned by this module's writer, by hand-assembled
bytes in tests/test_datasets.py and by the CRC / varint known answers -- parity with TF's writer is unpinned.
"""
import glob
import io
import os
import queue
import struct
import threading

import numpy as np

from .tf_checkpoint import _get_varint, _put_varint, crc32c, mask_crc

# keys read by _load_batch_pair_pose (trainer.py:540-542) and their per-example shapes; H, W are the dataset's image size
FEATURE_KEYS = ("image_raw_0", "image_raw_1", "label", "pose_peaks_0_rcv", "pose_peaks_1_rcv", "pose_mask_r4_0",
                "pose_mask_r4_1", "pose_mask_r6_0", "pose_mask_r6_1", "part_bbox_0", "part_bbox_1", "part_vis_0",
                "part_vis_1")
DATASETS = {"Market1501": (128, 64), "DeepFashion": (256, 256)}     # market1501.py:79-80, deepfashion.py:91-92


# ------------------------------------------------------------------------------------------ TFRecord framing
def read_records(path, verify=True):
    """Yields the payload bytes of every record of one TFRecord file."""
    with open(path, "rb") as fh:
        while True:
            head = fh.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError("%s: truncated record header" % path)
            (length,), (lcrc,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify and mask_crc(crc32c(head[:8])) != lcrc:
                raise ValueError("%s: corrupted record length" % path)
            data = fh.read(length)
            tail = fh.read(4)
            if len(data) < length or len(tail) < 4:
                raise ValueError("%s: truncated record" % path)
            if verify and mask_crc(crc32c(data)) != struct.unpack("<I", tail)[0]:
                raise ValueError("%s: corrupted record payload" % path)
            yield data


def write_records(path, payloads):
    with open(path, "wb") as fh:
        for data in payloads:
            head = struct.pack("<Q", len(data))
            fh.write(head)
            fh.write(struct.pack("<I", mask_crc(crc32c(head))))
            fh.write(data)
            fh.write(struct.pack("<I", mask_crc(crc32c(data))))


# ------------------------------------------------------------------------------------------ tf.Example
def _varints(buf):
    """All varints of a packed repeated field, vectorised (the pose masks are 8192-65536 values per example)."""
    b = np.frombuffer(buf, dtype=np.uint8)
    if b.size == 0:
        return np.zeros(0, np.int64)
    last = b < 0x80
    if last.all():                    # every value < 128: the {0,1} masks, small ints
        return b.astype(np.int64)
    ends = np.flatnonzero(last)
    starts = np.concatenate(([0], ends[:-1] + 1))
    pos = np.arange(b.size) - np.repeat(starts, ends - starts + 1)
    vals = (b & 0x7F).astype(np.uint64) << (7 * pos).astype(np.uint64)
    return np.add.reduceat(vals, starts).astype(np.int64)      # two's complement wrap gives negative int64 back


def parse_example(data):
    """tf.Example bytes -> dict key -> list of bytes | np.float32 array | np.int64 array."""
    out = {}
    pos, n = 0, len(data)
    while pos < n:                                 # Example: features = 1
        tag, pos = _get_varint(data, pos)
        ln, pos = _get_varint(data, pos)
        if tag != (1 << 3 | 2):
            pos += ln
            continue
        end = pos + ln
        while pos < end:                           # Features: repeated map entry feature = 1
            tag, pos = _get_varint(data, pos)
            eln, pos = _get_varint(data, pos)
            eend = pos + eln
            key, feat = None, b""
            while pos < eend:                      # map entry: key = 1, value = 2
                t2, pos = _get_varint(data, pos)
                l2, pos = _get_varint(data, pos)
                if t2 >> 3 == 1:
                    key = bytes(data[pos:pos + l2]).decode()
                elif t2 >> 3 == 2:
                    feat = data[pos:pos + l2]
                pos += l2
            if key is not None:
                out[key] = _parse_feature(feat)
    return out


def _parse_feature(feat):
    pos, n = 0, len(feat)
    if n == 0:
        return []
    tag, pos = _get_varint(feat, pos)
    ln, pos = _get_varint(feat, pos)
    kind, body = tag >> 3, feat[pos:pos + ln]
    p, m = 0, len(body)
    if kind == 1:                                  # BytesList: repeated bytes value = 1
        vals = []
        while p < m:
            _, p = _get_varint(body, p)
            l3, p = _get_varint(body, p)
            vals.append(bytes(body[p:p + l3]))
            p += l3
        return vals
    chunks = []
    while p < m:                                   # FloatList / Int64List: value = 1, packed (wire 2) or not
        t3, p = _get_varint(body, p)
        wt = t3 & 7
        if wt == 2:
            l3, p = _get_varint(body, p)
            chunks.append(np.frombuffer(body[p:p + l3], dtype="<f4") if kind == 2 else _varints(body[p:p + l3]))
            p += l3
        elif wt == 5:
            chunks.append(np.frombuffer(body[p:p + 4], dtype="<f4"))
            p += 4
        else:
            v, p = _get_varint(body, p)
            chunks.append(np.array([v], dtype=np.uint64).astype(np.int64))
    dt = np.float32 if kind == 2 else np.int64
    return np.concatenate(chunks).astype(dt) if chunks else np.zeros(0, dt)


def encode_example(features):
    """dict key -> bytes | list of bytes | float array | int array  ->  tf.Example bytes (packed lists, keys sorted the way
    protobuf's deterministic map serialisation orders them)."""
    feats = bytearray()
    for key in sorted(features):
        v = features[key]
        body = bytearray()
        if isinstance(v, (bytes, bytearray)) or (isinstance(v, list) and v and isinstance(v[0], (bytes, bytearray))):
            kind = 1
            for item in ([v] if isinstance(v, (bytes, bytearray)) else v):
                body.append(0x0A)
                _put_varint(body, len(item))
                body += item
        else:
            arr = np.asarray(v)
            packed = bytearray()
            if arr.dtype.kind == "f":
                kind = 2
                packed += arr.astype("<f4").tobytes()
            else:
                kind = 3
                for x in arr.reshape(-1).tolist():
                    _put_varint(packed, int(x))
            body.append(0x0A)
            _put_varint(body, len(packed))
            body += packed
        feat = bytearray([kind << 3 | 2])
        _put_varint(feat, len(body))
        feat += body
        kb = key.encode()
        entry = bytearray([0x0A])
        _put_varint(entry, len(kb))
        entry += kb
        entry.append(0x12)
        _put_varint(entry, len(feat))
        entry += feat
        feats.append(0x0A)
        _put_varint(feats, len(entry))
        feats += entry
    out = bytearray([0x0A])
    _put_varint(out, len(feats))
    out += feats
    return bytes(out)


# ------------------------------------------------------------------------------------------ decoding one pair
def decode_pair(example, img_h, img_w, part_num=37, keypoints=18):
    """One parsed tf.Example -> the per-sample tensors of _load_batch_pair_pose (trainer.py:544-551, 557-558)."""
    from PIL import Image

    def image(key):
        raw = example[key][0]
        im = np.asarray(Image.open(io.BytesIO(raw)).convert("RGB"), dtype=np.float32)
        if im.shape != (img_h, img_w, 3):
            raise ValueError("%s is %s, expected %dx%dx3" % (key, im.shape, img_h, img_w))
        return (im - 127.5) / 127.5                         # process_image(.., 127.5, 127.5)   utils.py:102-103

    out = {}
    for i, suffix in ((0, ""), (1, "_target")):
        out["x" + suffix] = image("image_raw_%d" % i)
        out["pose_rcv" + suffix] = np.asarray(example["pose_peaks_%d_rcv" % i], np.float32).reshape(keypoints, 3)
        out["mask" + suffix] = np.asarray(example["pose_mask_r6_%d" % i], np.float32).reshape(img_h, img_w, 1)
        out["mask_r4" + suffix] = np.asarray(example["pose_mask_r4_%d" % i], np.float32).reshape(img_h, img_w, 1)
        out["part_bbox" + suffix] = np.asarray(example["part_bbox_%d" % i], np.int64).reshape(part_num, 4)
        out["part_vis" + suffix] = np.asarray(example["part_vis_%d" % i], np.float32).reshape(part_num)
    out["label"] = np.int64(example["label"][0]) if "label" in example and len(example["label"]) else np.int64(0)
    return out


class TFRecordPairLoader:
    """`next_batch()` source for the trainers, standing where `_load_batch_pair_pose` + the TF queue runners stand
    (trainer.py:41-42, 105-110): a background thread reads `<data_name>_<split>_*.tfrecord` round-robin (like
    DatasetDataProvider without shuffling: `shuffle=True` draws from a 32-example reservoir as common_queue_capacity=32 /
    common_queue_min=8 do), decodes and stacks `batch_size` pairs, and hands them over through a bounded queue."""

    def __init__(self, dataset_dir, split="train", data_name="Market1501", batch_size=16, shuffle=True, seed=123,
                 prefetch=4, part_num=37, repeat=True):
        if data_name not in DATASETS:
            raise ValueError("data_name %r not in %s" % (data_name, sorted(DATASETS)))
        self.img_h, self.img_w = DATASETS[data_name]
        self.files = sorted(glob.glob(os.path.join(dataset_dir, "%s_%s_*.tfrecord" % (data_name, split))))
        if not self.files:
            raise IOError("no %s_%s_*.tfrecord under %s" % (data_name, split, dataset_dir))
        self.batch_size, self.shuffle, self.part_num, self.repeat = batch_size, shuffle, part_num, repeat
        self.rng = np.random.default_rng(seed)
        self.q = queue.Queue(maxsize=prefetch)
        self._stop = False
        self.thread = threading.Thread(target=self._produce, daemon=True)
        self.thread.start()

    def _examples(self):
        while True:
            for f in self.files:
                for rec in read_records(f):
                    yield rec
            if not self.repeat:
                return

    def _produce(self):
        try:
            pool, batch = [], []
            for rec in self._examples():
                if self._stop:
                    return
                if self.shuffle:
                    pool.append(rec)
                    if len(pool) < 32:
                        continue
                    rec = pool.pop(int(self.rng.integers(len(pool))))
                batch.append(decode_pair(parse_example(rec), self.img_h, self.img_w, self.part_num))
                if len(batch) == self.batch_size:
                    self.q.put({k: np.stack([b[k] for b in batch]) for k in batch[0]})
                    batch = []
            for rec in pool:       # drain the reservoir at the end of a non-repeating pass
                batch.append(decode_pair(parse_example(rec), self.img_h, self.img_w, self.part_num))
                if len(batch) == self.batch_size:
                    self.q.put({k: np.stack([b[k] for b in batch]) for k in batch[0]})
                    batch = []
            self.q.put(None)
        except Exception as e:      # surface reader errors in the consumer thread
            self.q.put(e)

    def next_batch(self):
        item = self.q.get()
        if isinstance(item, Exception):
            raise item
        if item is None:
            raise StopIteration("end of data")
        return item

    def close(self):
        self._stop = True


def get_split(split_name, dataset_dir, data_name="Market1501", batch_size=16, **kw):
    """datasets/market1501.py:50 / deepfashion.py get_split(): here it returns the ready batch source."""
    if split_name not in ("train", "test"):
        raise ValueError("split name %s was not recognized." % split_name)
    return TFRecordPairLoader(dataset_dir, split_name, data_name, batch_size, **kw)


def write_pair_records(path, pairs, quality=95, subsampling=-1):
    """Writes tf.Examples in the converters' schema (convert_market.py:_format_data): pairs = iterable of dicts with
    x / x_target uint8 [H,W,3], pose_rcv(_target) [18,3], mask(_target) and mask_r4(_target) [H,W], part_bbox(_target)
    [37,4], part_vis(_target) [37], label."""
    from PIL import Image
    payloads = []
    for p in pairs:
        f = {"image_format": b"jpg", "label": [int(p.get("label", 1))]}
        for i, s in ((0, ""), (1, "_target")):
            bio = io.BytesIO()
            Image.fromarray(np.asarray(p["x" + s], np.uint8)).save(bio, format="JPEG", quality=quality, subsampling=subsampling)
            f["image_raw_%d" % i] = bio.getvalue()
            f["image_name_%d" % i] = ("%s_%d.jpg" % (p.get("name", "img"), i)).encode()
            f["pose_peaks_%d_rcv" % i] = np.asarray(p["pose_rcv" + s], np.float32).reshape(-1)
            f["pose_mask_r6_%d" % i] = np.asarray(p["mask" + s]).astype(np.int64).reshape(-1)
            f["pose_mask_r4_%d" % i] = np.asarray(p.get("mask_r4" + s, p["mask" + s])).astype(np.int64).reshape(-1)
            f["part_bbox_%d" % i] = np.asarray(p["part_bbox" + s], np.int64).reshape(-1)
            f["part_vis_%d" % i] = np.asarray(p["part_vis" + s]).astype(np.int64).reshape(-1)
        h, w = np.asarray(p["x"]).shape[:2]
        f["image_height"], f["image_width"] = [h], [w]
        payloads.append(encode_example(f))
    write_records(path, payloads)
    return len(payloads)
