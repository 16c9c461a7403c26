"""Host side of the demo-compatible driver: the HDF5 pair reader and the PNG writer of demo.py without
h5py / torchvision / skimage.

  * ``read_h5_pair``  datasets/pix2pix.py:62-77: ``f['haze'][:]``, ``f['gt'][:]`` (H x W x 3 float64) -> two swapaxes ->
                      CHW.  The reader parses the subset of HDF5 that h5py writes for such files: version-0 superblock,
                      old-style root group (symbol table: B-tree v1 + local heap), version-1 object headers, contiguous
                      little-endian IEEE datasets.
  * ``write_png``     what ``vutils.save_image`` hands to PIL at demo.py:151, given the uint8 HWC bytes produced on the
                      GPU by ``fdgan_b200.metrics.save_image_u8``: an 8-bit RGB PNG (filter 0, zlib).
  * ``run_demo``      the loop of demo.py:118-151: read pairs ``<root>/<i>.h5``, run the generator, write ``<i>.png``.
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np
import torch

_SIG = b"\x89HDF\r\n\x1a\n"


class _H5:
    """Minimal read-only HDF5 (see module docstring).  Raises ValueError on anything outside the supported subset."""

    def __init__(self, buf: bytes):
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise ValueError("HDF5 superblock version %d is not supported (only version 0)" % buf[8])
        self.buf = buf
        self.O, self.L = buf[13], buf[14]
        if self.O != 8 or self.L != 8:
            raise ValueError("only 8-byte offsets / lengths are supported")
        self.base = self._u(24, 8)
        # root group symbol table entry starts after base / free-space / EOF / driver addresses
        entry = 24 + 4 * 8
        cache_type = self._u(entry + 16, 4)
        if cache_type != 1:
            raise ValueError("root group without cached symbol table")
        self.root_btree = self._u(entry + 24, 8)
        self.root_heap = self._u(entry + 32, 8)

    def _u(self, off: int, n: int) -> int:
        return int.from_bytes(self.buf[off:off + n], "little")

    # ---- group traversal ------------------------------------------------------------------------------------
    def _heap_data(self, heap_addr: int) -> int:
        a = self.base + heap_addr
        if self.buf[a:a + 4] != b"HEAP":
            raise ValueError("bad local heap")
        return self.base + self._u(a + 8 + 2 * 8, 8)          # signature, version + reserved (4), size, free list, data address

    def _name(self, heap_data: int, off: int) -> str:
        end = self.buf.index(b"\x00", heap_data + off)
        return self.buf[heap_data + off:end].decode()

    def _walk(self, btree_addr: int, heap_data: int, out: dict):
        a = self.base + btree_addr
        sig = self.buf[a:a + 4]
        if sig == b"TREE":
            if self.buf[a + 4] != 0:
                raise ValueError("unexpected B-tree node type")
            used = self._u(a + 6, 2)
            p = a + 8 + 2 * 8                                  # signature, type, level, entries, left / right siblings
            for i in range(used):
                child = self._u(p + 8 + i * 16, 8)             # key_i (8), child_i (8), ...
                self._walk(child, heap_data, out)
        elif sig == b"SNOD":
            n = self._u(a + 6, 2)
            p = a + 8
            for i in range(n):
                e = p + i * 40
                out[self._name(heap_data, self._u(e, 8))] = self._u(e + 8, 8)
        else:
            raise ValueError("bad group node")

    def members(self) -> dict:
        out = {}
        self._walk(self.root_btree, self._heap_data(self.root_heap), out)
        return out

    # ---- dataset --------------------------------------------------------------------------------------------
    def _messages(self, hdr_addr: int):
        a = self.base + hdr_addr
        if self.buf[a] != 1:
            raise ValueError("object header version %d is not supported" % self.buf[a])
        nmsgs = self._u(a + 2, 2)
        size = self._u(a + 8, 4)
        blocks = [(a + 16, size)]
        seen = 0
        while blocks and seen < nmsgs:
            p, left = blocks.pop(0)
            while left >= 8 and seen < nmsgs:
                mtype, msize = self._u(p, 2), self._u(p + 2, 2)
                data = p + 8
                seen += 1
                if mtype == 0x10:                              # continuation
                    blocks.append((self.base + self._u(data, 8), self._u(data + 8, 8)))
                else:
                    yield mtype, data, msize
                p += 8 + msize
                left -= 8 + msize

    def dataset(self, hdr_addr: int) -> np.ndarray:
        dims = dtype = layout = None
        for mtype, d, _size in self._messages(hdr_addr):
            if mtype == 0x1:                                   # dataspace
                ver, rank = self.buf[d], self.buf[d + 1]
                start = d + (8 if ver == 1 else 4)
                dims = [self._u(start + 8 * i, 8) for i in range(rank)]
            elif mtype == 0x3:                                 # datatype
                cls = self.buf[d] & 0x0F
                nbytes = self._u(d + 4, 4)
                big = self.buf[d + 1] & 1
                if cls == 1:
                    dtype = np.dtype((">" if big else "<") + "f%d" % nbytes)
                elif cls == 0:
                    signed = (self.buf[d + 1] >> 3) & 1
                    dtype = np.dtype((">" if big else "<") + ("i" if signed else "u") + "%d" % nbytes)
                else:
                    raise ValueError("unsupported datatype class %d" % cls)
            elif mtype == 0x8:                                 # data layout
                if self.buf[d] != 3 or self.buf[d + 1] != 1:
                    raise ValueError("only contiguous version-3 layouts are supported")
                layout = (self._u(d + 2, 8), self._u(d + 10, 8))
        if dims is None or dtype is None or layout is None:
            raise ValueError("incomplete dataset header")
        addr, nbytes = layout
        count = int(np.prod(dims)) if dims else 1
        if count * dtype.itemsize > nbytes:
            raise ValueError("dataset larger than its storage")
        return np.frombuffer(self.buf, dtype, count, self.base + addr).reshape(dims)


def read_h5(path: str, names=("haze", "gt")) -> dict:
    with open(path, "rb") as f:
        h5 = _H5(f.read())
    members = h5.members()
    missing = [n for n in names if n not in members]
    if missing:
        raise KeyError("datasets %s not in %s (has %s)" % (missing, path, sorted(members)))
    return {n: h5.dataset(members[n]) for n in names}


def read_h5_pair(path: str):
    """-> (haze, gt) as float32 CHW tensors, exactly the arrays datasets/pix2pix.py:62-77 hands to the DataLoader
    (before the ``.float()`` of demo.py:126)."""
    d = read_h5(path)
    to_chw = lambda a: np.ascontiguousarray(np.swapaxes(np.swapaxes(a, 0, 2), 1, 2))
    return torch.from_numpy(to_chw(d["haze"])).float(), torch.from_numpy(to_chw(d["gt"])).float()


def write_png(path: str, img_u8) -> None:
    """img_u8: uint8 [H,W,3] (torch tensor on any device, or numpy) -> 8-bit RGB PNG."""
    a = img_u8.detach().cpu().numpy() if torch.is_tensor(img_u8) else np.asarray(img_u8)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("write_png expects uint8 [H,W,3]")
    h, w, _ = a.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), np.ascontiguousarray(a).reshape(h, w * 3)], axis=1).tobytes()   # filter type 0 per scanline

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def run_demo(netG, data_root: str, out_dir: str, count: int, first_index: int = 0, device="cuda"):
    """demo.py:118-151 for ``count`` samples ``<data_root>/<i>.h5`` (valBatchSize 1): generator forward in the module's
    current mode (the reference never calls .eval(): README.md:38), min-max normalised PNG ``<out_dir>/<k>.png``.
    Returns the list of (uint8 output, uint8 ground truth) CUDA tensors for metric evaluation."""
    from . import metrics
    os.makedirs(out_dir, exist_ok=True)
    pairs = []
    with torch.no_grad():
        for k in range(count):
            haze, gt = read_h5_pair(os.path.join(data_root, "%d.h5" % (first_index + k)))
            y = netG(haze.unsqueeze(0).to(device))
            out_u8 = metrics.save_image_u8(y[0])
            write_png(os.path.join(out_dir, "%d.png" % k), out_u8)
            pairs.append((out_u8, metrics.save_image_u8(gt.to(device))))
    return pairs
