"""Test helper: minimal OpenEXR reader, independent of gpurt_write_exr."""
import numpy as np


def read_exr(path):
    """Minimal OpenEXR reader written from the file-layout document, independent of the writer: single part, scanlines,
    no compression, FLOAT channels.  Returns (attributes, {channel name: (h, w) float32})."""
    import struct
    b = open(path, "rb").read()
    magic, version = struct.unpack_from("<iI", b, 0)
    assert magic == 20000630 and version & 0xFF == 2 and version >> 8 == 0, "magic / version / flag bits"
    pos, attrs = 8, {}

    def cstr(p):
        e = b.index(b"\0", p)
        return b[p:e].decode(), e + 1
    while b[pos] != 0:
        name, pos = cstr(pos)
        typ, pos = cstr(pos)
        size, = struct.unpack_from("<i", b, pos)
        attrs[name] = (typ, b[pos + 4:pos + 4 + size])
        pos += 4 + size
    pos += 1
    for need in ("channels", "compression", "dataWindow", "displayWindow", "lineOrder", "pixelAspectRatio",
                 "screenWindowCenter", "screenWindowWidth"):
        assert need in attrs, f"required attribute {need} missing"
    assert attrs["compression"] == ("compression", b"\0") and attrs["lineOrder"] == ("lineOrder", b"\0")
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    chans, c, p = [], attrs["channels"][1], 0
    while c[p] != 0:
        e = c.index(b"\0", p)
        name = c[p:e].decode()
        ptype, plinear, xs, ys = struct.unpack_from("<iB3xii", c, e + 1)
        assert ptype == 2 and xs == 1 and ys == 1
        chans.append(name)
        p = e + 1 + 16
    assert p + 1 == len(c) and chans == sorted(chans), "channel list must be sorted and fill its attribute"
    offsets = struct.unpack_from(f"<{h}Q", b, pos)
    out = {n: np.zeros((h, w), np.float32) for n in chans}
    for i, off in enumerate(offsets):
        y, nbytes = struct.unpack_from("<ii", b, off)
        assert y == y0 + i and nbytes == 4 * w * len(chans)
        row = np.frombuffer(b, "<f4", w * len(chans), off + 8).reshape(len(chans), w)
        for k, n in enumerate(chans):
            out[n][i] = row[k]
    assert offsets[-1] + 8 + 4 * w * len(chans) == len(b), "nothing after the last scanline"
    return attrs, out
