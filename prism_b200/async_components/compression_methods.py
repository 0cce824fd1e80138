"""Message envelope of the Redis transport: ``msgpack((compression_type, msgpack(payload) [compressed]))``.

Wire-compatible with the reference's ``MessageSerializer`` (prism/async_components/compression_methods.py:37-81):
same two-level msgpack envelope, same compression tags ("NONE", "LZ4" = an LZ4 *frame*), same size threshold.
Two differences, both on the host side of the wire only:

* numeric payloads (every payload this transport carries is one flat list of numbers) are encoded / decoded by the
  native codec in csrc/wire.cu (``pack_numbers`` / ``unpack_numbers``) instead of element by element in Python;
* the LZ4 frame codec is looked up when first needed: the ``lz4`` module, else pyarrow's bundled liblz4 (same frame
  format).  A sender with neither falls back to the "NONE" tag (the envelope is self-describing, so any receiver
  reads it); a receiver with neither raises on an "LZ4" message -- never a silent skip.
"""
import ctypes as C

import msgpack
import numpy as np

from .. import _lib

MIN_SIZE_TO_COMPRESS = 1024

LZ4 = True
NONE = False

_DTYPE_CODES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int64): 2, np.dtype(np.bool_): 3}


class _ArrowLZ4Frame(object):
    """LZ4 *frame* format through pyarrow's bundled liblz4 -- the same container ``lz4.frame`` reads and writes."""

    def __init__(self, pa):
        self._pa = pa

    def compress(self, data):
        return self._pa.Codec("lz4").compress(data, asbytes=True)

    def decompress(self, data):
        # streaming reader: python-lz4 frames carry linked blocks and an optional content size; both are handled
        return self._pa.CompressedInputStream(self._pa.BufferReader(data), "lz4").read()


_LZ4_PROVIDER = []


def _lz4_frame():
    """The LZ4 frame codec: the ``lz4`` module when installed (what the reference uses), else pyarrow's."""
    if not _LZ4_PROVIDER:
        provider = None
        try:
            import lz4.frame
            provider = lz4.frame
        except ImportError:
            try:
                import pyarrow
                if pyarrow.Codec.is_available("lz4"):
                    provider = _ArrowLZ4Frame(pyarrow)
            except ImportError:
                pass
        _LZ4_PROVIDER.append(provider)
    return _LZ4_PROVIDER[0]


class NullMessageCompressor(object):
    compression_type = "NONE"

    @staticmethod
    def compress(data):
        return data

    @staticmethod
    def decompress(data):
        return data


class LZ4MessageCompressor(object):
    compression_type = "LZ4"

    @staticmethod
    def available():
        return _lz4_frame() is not None

    @staticmethod
    def compress(data):
        return _lz4_frame().compress(data)

    @staticmethod
    def decompress(data):
        frame = _lz4_frame()
        if frame is None:
            raise _lib.PbError("received an LZ4-compressed message but neither lz4 nor pyarrow is installed")
        return frame.decompress(data)


def pack_numbers(segments):
    """msgpack bytes of ONE flat list made of ``segments`` (numpy arrays: float32 / float64 / int64 / bool, or Python
    ints) -- byte-identical to ``msgpack.packb`` of the concatenated Python list (csrc/wire.cu)."""
    lib = _lib.load()
    arrays = []
    for seg in segments:
        a = np.ascontiguousarray(seg)
        if a.dtype not in _DTYPE_CODES:
            if np.issubdtype(a.dtype, np.integer):
                a = a.astype(np.int64)
            elif np.issubdtype(a.dtype, np.floating):
                a = a.astype(np.float64)
            else:
                raise TypeError("cannot put dtype %s on the wire" % a.dtype)
        arrays.append(a.reshape(-1))
    total = sum(a.size for a in arrays)
    out = np.empty(5 + 9 * total, dtype=np.uint8)
    written = C.c_longlong(0)
    _lib.check(lib.pb_wire_array_header(total, out.ctypes.data, C.byref(written)), "pb_wire_array_header")
    pos = written.value
    for a in arrays:
        if a.size == 0:
            continue
        _lib.check(lib.pb_wire_pack_numbers(a.ctypes.data, _DTYPE_CODES[a.dtype], a.size, out.ctypes.data + pos,
                                            out.size - pos, C.byref(written)), "pb_wire_pack_numbers")
        pos += written.value
    return out[:pos].tobytes()


def unpack_numbers(data):
    """Inverse of ``pack_numbers`` for any msgpack array of ints / floats / bools: a float64 numpy array."""
    lib = _lib.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    n = C.c_longlong(0)
    if buf.size >= 1:
        head = int(buf[0])
        if (head & 0xf0) == 0x90:
            n.value = head & 0x0f
        elif head == 0xdc and buf.size >= 3:
            n.value = int.from_bytes(bytes(buf[1:3]), "big")
        elif head == 0xdd and buf.size >= 5:
            n.value = int.from_bytes(bytes(buf[1:5]), "big")
    if n.value > buf.size:              # every element takes at least one byte: refuse to allocate from a bad header
        raise _lib.PbError("malformed numeric payload (array header claims %d elements in %d bytes)" % (n.value, buf.size))
    out = np.empty(max(n.value, 1), dtype=np.float64)
    _lib.check(lib.pb_wire_unpack_numbers(buf.ctypes.data, buf.size, out.ctypes.data, out.size, C.byref(n)),
               "pb_wire_unpack_numbers")
    return out[:n.value]


class MessageSerializer(object):
    _compressors = {
        NullMessageCompressor.compression_type: NullMessageCompressor,
        LZ4MessageCompressor.compression_type: LZ4MessageCompressor,
    }

    def __init__(self, compression_type=LZ4MessageCompressor.compression_type, min_size_to_compress=MIN_SIZE_TO_COMPRESS):
        self._compression_type = compression_type.upper() if compression_type else None
        self._min_size_to_compress = min_size_to_compress

    @classmethod
    def register_compressor(cls, compressor):
        MessageSerializer._compressors[compressor.compression_type.upper()] = compressor

    # ---- envelope ---------------------------------------------------------------------------------------------
    def _wrap(self, data):
        compression_type = "NONE"
        if self._compression_type and self._compression_type != "NONE" and self._min_size_to_compress <= len(data):
            compressor = MessageSerializer._compressors[self._compression_type]
            if getattr(compressor, "available", lambda: True)():
                compression_type = self._compression_type
                data = compressor.compress(data)
        return msgpack.packb((compression_type, data))

    def _unwrap(self, data):
        compression_type, data = msgpack.unpackb(data)
        if isinstance(compression_type, bytes):
            compression_type = compression_type.decode()
        try:
            compressor = MessageSerializer._compressors[compression_type]
        except KeyError:
            raise ValueError("Received message with unknown compression type '%s'. Supported types are %s"
                             % (compression_type, ",".join(MessageSerializer._compressors.keys())))
        return compressor.decompress(data)

    # ---- reference API ----------------------------------------------------------------------------------------
    def pack(self, data):
        if data is None:
            return None
        return self._wrap(msgpack.packb(data))

    def unpack(self, data):
        if data is None:
            return None
        return msgpack.unpackb(self._unwrap(data))

    # ---- numeric fast path ------------------------------------------------------------------------------------
    def pack_numbers(self, segments):
        return self._wrap(pack_numbers(segments))

    def unpack_numbers(self, data):
        if data is None:
            return None
        return unpack_numbers(self._unwrap(data))
