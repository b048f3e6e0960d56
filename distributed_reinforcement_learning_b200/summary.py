"""Stand-in for ``tensorboardX.SummaryWriter`` as the reference's launchers use it (``SummaryWriter(logdir)`` +
``add_scalar(tag, value, step)``; train_impala.py:91,109-113, train_apex.py:92,143-144, train_r2d2.py:90,157-158):
writes real TensorBoard event files (``events.out.tfevents.*``) without TensorFlow, tensorboardX or protobuf.

File format (TFRecord): per record  uint64 length | uint32 masked_crc32c(length) | bytes | uint32 masked_crc32c(bytes);
the payload is a serialized ``tensorflow.Event`` protobuf, hand-encoded here:
  Event   { double wall_time = 1; int64 step = 2; string file_version = 3; Summary summary = 5; }
  Summary { repeated Value value = 1; }   Value { string tag = 1; float simple_value = 2; }
"""
import os
import socket
import struct
import time

_CRC_TABLE = []


def _crc_table():
    if not _CRC_TABLE:
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1     # CRC-32C (Castagnoli), reflected
            _CRC_TABLE.append(c)
    return _CRC_TABLE


def crc32c(data):
    t = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n):
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field_bytes(num, payload):
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def encode_scalar_event(tag, value, step, wall_time):
    val = _field_bytes(1, tag.encode("utf-8")) + _varint((2 << 3) | 5) + struct.pack("<f", float(value))
    summary = _field_bytes(1, val)
    return (_varint((1 << 3) | 1) + struct.pack("<d", wall_time) + _varint((2 << 3) | 0) + _varint(int(step)) +
            _field_bytes(5, summary))


def encode_version_event(wall_time):
    return _varint((1 << 3) | 1) + struct.pack("<d", wall_time) + _field_bytes(3, b"brain.Event:2")


def _record(payload):
    head = struct.pack("<Q", len(payload))
    return head + struct.pack("<I", masked_crc(head)) + payload + struct.pack("<I", masked_crc(payload))


class SummaryWriter:
    def __init__(self, logdir=None, **_ignored):
        self.logdir = logdir or os.path.join("runs", time.strftime("%b%d_%H-%M-%S") + "_" + socket.gethostname())
        os.makedirs(self.logdir, exist_ok=True)
        name = "events.out.tfevents.%010d.%s.%d" % (int(time.time()), socket.gethostname(), os.getpid())
        self.path = os.path.join(self.logdir, name)
        self._f = open(self.path, "ab")
        self._f.write(_record(encode_version_event(time.time())))
        self._f.flush()

    def add_scalar(self, tag, scalar_value, global_step=None, walltime=None):
        wt = time.time() if walltime is None else walltime
        self._f.write(_record(encode_scalar_event(tag, scalar_value, 0 if global_step is None else global_step, wt)))

    def flush(self):
        self._f.flush()

    def close(self):
        if self._f and not self._f.closed:
            self._f.flush()
            self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def read_scalars(path):
    """Test aid: decode an event file written above -> [(tag, value, step, wall_time)], checking every CRC."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    off = 0

    def varint(buf, i):
        n = shift = 0
        while True:
            b = buf[i]
            i += 1
            n |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                return n, i

    def fields(buf):
        i = 0
        while i < len(buf):
            key, i = varint(buf, i)
            num, wt = key >> 3, key & 7
            if wt == 0:
                v, i = varint(buf, i)
            elif wt == 1:
                v, i = buf[i:i + 8], i + 8
            elif wt == 5:
                v, i = buf[i:i + 4], i + 4
            elif wt == 2:
                ln, i = varint(buf, i)
                v, i = buf[i:i + ln], i + ln
            else:
                raise ValueError("wire type %d" % wt)
            yield num, wt, v
    while off < len(data):
        head = data[off:off + 8]
        (ln,) = struct.unpack("<Q", head)
        (c1,) = struct.unpack("<I", data[off + 8:off + 12])
        payload = data[off + 12:off + 12 + ln]
        (c2,) = struct.unpack("<I", data[off + 12 + ln:off + 16 + ln])
        if c1 != masked_crc(head) or c2 != masked_crc(payload):
            raise ValueError("CRC mismatch at offset %d" % off)
        off += 16 + ln
        ev = {num: v for num, _, v in fields(payload)}
        if 5 not in ev:
            continue
        wall = struct.unpack("<d", ev[1])[0]
        step = ev.get(2, 0)
        for num, _, val in fields(ev[5]):
            if num == 1:
                vv = {n2: v2 for n2, _, v2 in fields(val)}
                out.append((vv[1].decode(), struct.unpack("<f", vv[2])[0], step, wall))
    return out
