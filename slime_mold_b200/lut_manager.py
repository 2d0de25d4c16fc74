"""LUTs -- mirror of /root/reference/src/lut_manager.rs (LutData, LutManager).

The reference embeds 109 of its `LUTs/*.lut` files at compile time (lut_manager.rs:36-146);
a .lut file is 768 bytes: 256 red, 256 green, 256 blue (lut_manager.rs:162-186).  This mirror
reads the same files from a directory the caller names (the reference's `LUTs/` checkout) and
ships two procedural tables for hosts that have none: "gray" and "gray_r" (the reference's
default LUT, MATPLOTLIB_bone_r, is a data file of the reference and is not copied here).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np


class LutData:
    """lut_manager.rs:4-18"""

    def __init__(self, name: str, red, green, blue):
        self.name = name
        self.red = np.asarray(red, dtype=np.uint8).copy()
        self.green = np.asarray(green, dtype=np.uint8).copy()
        self.blue = np.asarray(blue, dtype=np.uint8).copy()
        if not (self.red.size == self.green.size == self.blue.size == 256):
            raise ValueError("each LUT component must hold 256 bytes")

    def reverse(self) -> None:
        """lut_manager.rs:13-17"""
        self.red = self.red[::-1].copy()
        self.green = self.green[::-1].copy()
        self.blue = self.blue[::-1].copy()

    def combined(self) -> np.ndarray:
        """red ++ green ++ blue, the 768 bytes main.rs:330-334 hands to the display shader."""
        return np.concatenate([self.red, self.green, self.blue]).astype(np.uint8)


class LutManager:
    """lut_manager.rs:149-186; `directory` stands where the reference's embedded table stands."""

    def __init__(self, directory: Optional[str] = None):
        self._files: Dict[str, str] = {}
        if directory:
            for fn in os.listdir(directory):
                if fn.endswith(".lut"):
                    self._files[fn[:-4]] = os.path.join(directory, fn)
        ramp = np.arange(256, dtype=np.uint8)
        self._builtin = {"gray": LutData("gray", ramp, ramp, ramp), "gray_r": LutData("gray_r", ramp[::-1], ramp[::-1], ramp[::-1])}

    def get_available_luts(self) -> List[str]:
        """sorted names, lut_manager.rs:154-158"""
        return sorted(set(self._files) | set(self._builtin))

    def load_lut(self, name: str) -> LutData:
        """lut_manager.rs:160-186: NotFound for an unknown name, InvalidData unless the file is 768 bytes."""
        if name in self._files:
            buf = np.fromfile(self._files[name], dtype=np.uint8)
            if buf.size != 768:
                raise ValueError("Invalid LUT file size")
            return LutData(name, buf[0:256], buf[256:512], buf[512:768])
        if name in self._builtin:
            b = self._builtin[name]
            return LutData(b.name, b.red, b.green, b.blue)
        raise FileNotFoundError(f"LUT '{name}' not found")


def write_png(path: str, rgba: np.ndarray) -> None:
    """Minimal RGBA8 PNG encoder (zlib + CRC from the standard library) for headless frame dumps."""
    import struct
    import zlib
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w, c = a.shape
    assert c == 4
    raw = np.empty((h, 1 + w * 4), np.uint8)
    raw[:, 0] = 0                      # filter type 0 on every scanline
    raw[:, 1:] = a.reshape(h, w * 4)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)))
        f.write(chunk(b"IEND", b""))
