"""The output side of the boundary: the film as the reference writes it.

`write_exr` (src/main.rs:42-53) hands the row-major `Vec<Color>` to `exr::write_rgb_file`: a
single-part scan-line OpenEXR file with three 32-bit FLOAT channels.  `write_exr` here emits the
same image as an UNCOMPRESSED scan-line file (any EXR reader opens it; pixel values are
bit-identical to the film), and `exr_file_name` reproduces the reference's naming
(src/main.rs:238-243): "{scene}-{integrator}-{spp}spp.exr".  `read_exr` reads such files back
(tests).  `write_png` mirrors `write_image` (:28-40) for 8-bit previews.
"""
import struct

import numpy as np


def exr_file_name(scene_name, integrator, msaa):
    return f"{scene_name}-{integrator}-{msaa * msaa}spp.exr"


def _attr(name, typ, payload):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def write_exr(path, film):
    """film: float32 [H, W, 3] (row 0 = top), exactly what pbrs_render returns."""
    film = np.ascontiguousarray(film, dtype=np.float32)
    h, w, c = film.shape
    assert c == 3
    # channel list, alphabetical as the format requires: B, G, R; pixel type 2 = FLOAT
    chans = b"".join(n + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) for n in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = b"".join([
        _attr("channels", "chlist", chans),
        _attr("compression", "compression", b"\0"),  # NO_COMPRESSION
        _attr("dataWindow", "box2i", box),
        _attr("displayWindow", "box2i", box),
        _attr("lineOrder", "lineOrder", b"\0"),      # INCREASING_Y
        _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)),
        _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)),
        _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)),
    ]) + b"\0"
    magic = struct.pack("<iI", 20000630, 2)  # version 2, single-part scan line
    line_bytes = 3 * w * 4
    table_pos = len(magic) + len(header)
    first = table_pos + 8 * h
    offsets = np.arange(h, dtype=np.uint64) * np.uint64(8 + line_bytes) + np.uint64(first)
    # each scan line block: y, byte count, then the channels one after another (B row, G row, R row)
    body = np.empty((h, 8 + line_bytes), np.uint8)
    body[:, 0:4] = np.arange(h, dtype="<i4").view(np.uint8).reshape(h, 4)
    body[:, 4:8] = np.frombuffer(struct.pack("<i", line_bytes), np.uint8)
    planes = np.ascontiguousarray(film[:, :, ::-1].transpose(0, 2, 1))  # [H, (B,G,R), W]
    body[:, 8:] = planes.view(np.uint8).reshape(h, line_bytes)
    with open(path, "wb") as f:
        f.write(magic)
        f.write(header)
        f.write(offsets.astype("<u8").tobytes())
        f.write(body.tobytes())


def read_exr(path):
    """Reads back an uncompressed scan-line FLOAT RGB file written by write_exr -> float32 [H, W, 3]."""
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<iI", data, 0)
    assert magic == 20000630 and (version & 0xFF) == 2
    pos, attrs = 8, {}
    while data[pos] != 0:
        end = data.index(b"\0", pos); name = data[pos:end].decode(); pos = end + 1
        end = data.index(b"\0", pos); typ = data[pos:end].decode(); pos = end + 1
        (size,) = struct.unpack_from("<i", data, pos); pos += 4
        attrs[name] = (typ, data[pos:pos + size]); pos += size
    pos += 1
    assert attrs["compression"][1] == b"\0", "only uncompressed files"
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    names, cp, ch = [], 0, attrs["channels"][1]
    while ch[cp] != 0:
        end = ch.index(b"\0", cp); names.append(ch[cp:end].decode()); cp = end + 1 + 16
    offsets = np.frombuffer(data, "<u8", h, pos)
    film = np.zeros((h, w, 3), np.float32)
    for off in offsets:
        y, nbytes = struct.unpack_from("<ii", data, int(off))
        row = np.frombuffer(data, "<f4", len(names) * w, int(off) + 8).reshape(len(names), w)
        for k, n in enumerate(names):
            film[y - y0, :, "RGB".index(n)] = row[k]
    return film


def write_png(path, film, gamma=2.2):
    """8-bit preview (src/main.rs:28-40 write_image)."""
    from PIL import Image
    img = np.clip(np.nan_to_num(film), 0.0, 1.0) ** (1.0 / gamma)
    Image.fromarray((img * 255.0 + 0.5).astype(np.uint8)).save(path)
