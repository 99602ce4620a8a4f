"""A minimal protobuf WIRE-FORMAT encoder for the test fixtures of the `.caffemodel` path: the test writes the bytes a
protobuf library would write for src/caffe/proto/caffe.proto messages (field numbers cited in
caffe_escoin_b200/host/escort_caffemodel.hpp) without needing protobuf or a compiled caffe_pb2."""
import struct


def varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def tag(num, wt):
    return varint((num << 3) | wt)


def f_uint(num, v):
    return tag(num, 0) + varint(v)


def f_bytes(num, b):
    if isinstance(b, str):
        b = b.encode()
    return tag(num, 2) + varint(len(b)) + b


def f_packed_float(num, arr):
    return f_bytes(num, struct.pack("<%df" % len(arr), *arr))


def f_packed_double(num, arr):
    return f_bytes(num, struct.pack("<%dd" % len(arr), *arr))


def f_float_unpacked(num, arr):
    return b"".join(tag(num, 5) + struct.pack("<f", v) for v in arr)


def blob_shape(data, shape):
    """BlobProto with BlobShape shape = 7 { dim = 1 [packed] } and packed float data = 5."""
    dims = b"".join(varint(d) for d in shape)
    return f_bytes(7, f_bytes(1, dims)) + f_packed_float(5, data)


def blob_legacy(data, nchw, unpacked=False):
    """BlobProto with the deprecated num / channels / height / width = 1..4."""
    head = b"".join(f_uint(i + 1, d) for i, d in enumerate(nchw))
    return head + (f_float_unpacked(5, data) if unpacked else f_packed_float(5, data))


def blob_double(data, shape):
    dims = b"".join(varint(d) for d in shape)
    return f_bytes(7, f_bytes(1, dims)) + f_packed_double(8, data)


def conv_param(num_output, kernel, stride=1, pad=0, group=1, bias_term=True, dilation=1, hw=None):
    p = f_uint(1, num_output) + f_uint(2, int(bias_term)) + f_uint(5, group)
    if hw is None:
        p += f_uint(3, pad) + f_uint(4, kernel) + f_uint(6, stride)        # repeated uint32, one element each
    else:                                                                 # the 2-D-only fields
        p += f_uint(11, hw["kernel_h"]) + f_uint(12, hw["kernel_w"]) + f_uint(9, hw["pad_h"]) + f_uint(10, hw["pad_w"])
        p += f_uint(13, hw["stride_h"]) + f_uint(14, hw["stride_w"])
    if dilation != 1:
        p += f_uint(18, dilation)
    return p


def layer_v2(name, type_, blobs, conv=None, ip=None, extra=b""):
    """LayerParameter: name = 1, type = 2, blobs = 7, convolution_param = 106, inner_product_param = 117."""
    b = f_bytes(1, name) + f_bytes(2, type_) + f_bytes(3, "bottom_of_" + name) + extra
    if conv is not None:
        b += f_bytes(106, conv)
    if ip is not None:
        b += f_bytes(117, ip)
    for bl in blobs:
        b += f_bytes(7, bl)
    return f_bytes(100, b)


def layer_v1(name, type_enum, blobs, conv=None):
    """V1LayerParameter: name = 4, type = 5 (enum), blobs = 6, convolution_param = 10."""
    b = f_bytes(4, name) + f_uint(5, type_enum)
    if conv is not None:
        b += f_bytes(10, conv)
    for bl in blobs:
        b += f_bytes(6, bl)
    return f_bytes(2, b)


def net(name, layers, extra=b""):
    return f_bytes(1, name) + extra + b"".join(layers)
