"""h5lite -- a self-contained reader / writer for the subset of HDF5 the AtlasPatch container uses, behind the h5py calls
the reference makes (so `atlaspatch_b200.storage` -- and the reference's own `H5PatchWriter`, unmodified -- run where neither
h5py nor libhdf5 exists, which is the case in this build image and on the GPU boxes).

reference: atlas_patch/utils/h5.py:11-97 (File "w", create_dataset(shape, maxshape, chunks, dtype), Dataset.resize / slicing,
File.attrs, Dataset.attrs), services/storage.py:250-385 (File "a", require_group, `in`, `del`, Group.move), utils/features.py:37-60
and orchestration/runner.py:77-88 (File "r", attrs.get, group iteration, Dataset.shape).

File format written (HDF5 File Format Specification v3.0, the "earliest" layout libhdf5 / h5py emit by default):
  superblock v0 (8-byte offsets / lengths) * old-style groups: v1 object header + symbol-table message -> B-tree v1 (group
  nodes, K = 16) + local heap + SNOD symbol-table nodes (leaf K = 4: 8 entries each) * datasets: v1 object header with
  dataspace v1 (max dims; H5S_UNLIMITED), datatype v1 (int / float little endian, fixed NULL-padded strings), fill value v2,
  data layout v3 CHUNKED -> B-tree v1 of raw chunks (K = 32: 64 entries per node, any depth), no filters * attributes v1
  (scalars / arrays of int / float, fixed strings, variable-length UTF-8 strings in global heap collections "GCOL").
The reader understands the same subset plus what old files add: superblock v0 / v1 with a base address (user block), object
header continuation blocks, contiguous and compact layouts (v1-v3), deflate / shuffle filters.  New-style groups (superblock
v2+, "OHDR" headers, fractal heaps) are rejected with a clear error.

PARITY STATUS: the writer's bytes follow the published specification and are read back by the independently written reader
below (tests/test_h5lite.py), and the reader is checked against a file written by the real HDF5 library that ships in this image
(scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat: superblock, B-tree, heap, SNOD, v1 headers, attributes).  No libhdf5 can
open files here, so byte-level acceptance by h5py remains UNPINNED until the test suite runs where h5py is installed
(tests/test_h5lite.py::test_h5py_reads_h5lite_file is skipped without it).

Semantics differ from h5py in one deliberate way: a file is parsed into memory on open and serialised on close / flush (files of
this container are a few hundred MB at most); datasets are numpy arrays in between.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Any, Iterator

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b"\x89HDF\r\n\x1a\n"
GROUP_LEAF_K, GROUP_INTERNAL_K, ISTORE_K = 4, 16, 32


# =====================================================================================================================
# in-memory tree (the h5py-like API)
# =====================================================================================================================
class AttributeManager:
    def __init__(self, owner):
        self._d: dict[str, Any] = {}
        self._owner = owner

    def _touch(self):
        f = self._owner._file
        if f is not None:
            f._require_writable()
            f._dirty = True

    def __setitem__(self, k: str, v: Any) -> None:
        self._touch()
        self._d[str(k)] = _normalise_attr(v)

    def __getitem__(self, k: str) -> Any:
        return self._d[k]

    def __delitem__(self, k: str) -> None:
        self._touch()
        del self._d[k]

    def __contains__(self, k) -> bool:
        return k in self._d

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def get(self, k, default=None):
        return self._d.get(k, default)

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def values(self):
        return self._d.values()

    def update(self, other):
        for k, v in dict(other).items():
            self[k] = v


def _normalise_attr(v: Any) -> Any:
    """What h5py would store: Python int -> int64, float -> float64, str -> variable-length UTF-8, bytes -> fixed string."""
    if isinstance(v, (bool, np.bool_)):
        return np.int8(1 if v else 0)          # h5py stores an enum; the container never writes booleans
    if isinstance(v, str):
        return v
    if isinstance(v, (bytes, np.bytes_)):
        return np.bytes_(v)
    if isinstance(v, int):
        return np.int64(v)
    if isinstance(v, float):
        return np.float64(v)
    if isinstance(v, np.generic):
        return v
    a = np.asarray(v)
    if a.dtype.kind in "iuf":
        return a.copy()
    if a.dtype.kind == "S":
        return a.copy()
    if a.dtype.kind in "UO":
        if a.ndim == 0:
            return str(a.item())
        raise TypeError("h5lite: arrays of Python strings are not supported as attributes")
    raise TypeError(f"h5lite: unsupported attribute value of type {type(v)!r}")


class _Node:
    def __init__(self, file, name: str):
        self._file = file
        self.name = name
        self.attrs = AttributeManager(self)


class Dataset(_Node):
    def __init__(self, file, name, data: np.ndarray, maxshape, chunks):
        super().__init__(file, name)
        self._buf = data                      # capacity >= shape along axis 0 (appends grow it geometrically)
        self._shape = tuple(data.shape)
        self.maxshape = tuple(maxshape)
        self.chunks = tuple(chunks) if chunks is not None else None

    @property
    def _data(self) -> np.ndarray:
        return self._buf[:self._shape[0]] if self._shape else self._buf

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def ndim(self):
        return self._data.ndim

    @property
    def size(self):
        return int(self._data.size)

    def __len__(self):
        return int(self._data.shape[0])

    def resize(self, size, axis=None) -> None:
        self._file._require_writable()
        if axis is not None:
            new = list(self.shape)
            new[int(axis)] = int(size)
        else:
            new = [int(s) for s in (size if isinstance(size, (tuple, list)) else (size,))]
        if len(new) != self._data.ndim:
            raise TypeError("h5lite: resize changes the rank")
        for n, m in zip(new, self.maxshape):
            if m is not None and n > m:
                raise ValueError(f"h5lite: unable to resize dataset beyond maxshape {self.maxshape}")
        if self.chunks is None:
            raise TypeError("h5lite: only chunked datasets can be resized")
        old = self._shape
        if tuple(new[1:]) == tuple(old[1:]) and new[0] <= self._buf.shape[0]:
            if new[0] < old[0]:
                self._buf[new[0]:old[0]] = 0   # rows given up read back as the fill value if the dataset grows again
        else:
            cap = list(new)
            if tuple(new[1:]) == tuple(old[1:]):
                cap[0] = max(new[0], 2 * self._buf.shape[0], 16)
            out = np.zeros(tuple(cap), dtype=self._buf.dtype)
            sl = tuple(slice(0, min(a, b)) for a, b in zip(new, old))
            out[sl] = self._data[sl]
            self._buf = out
        self._shape = tuple(new)
        self._file._dirty = True

    def __getitem__(self, idx):
        out = self._data[idx]
        return out.copy() if isinstance(out, np.ndarray) else out

    def __setitem__(self, idx, val) -> None:
        self._file._require_writable()
        self._data[idx] = val
        self._file._dirty = True

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def __repr__(self):
        return f'<h5lite dataset "{self.name}": shape {self.shape}, type "{self.dtype.str}">'


class Group(_Node):
    def __init__(self, file, name):
        super().__init__(file, name)
        self._children: dict[str, _Node] = {}

    # ---- navigation ----
    def _walk(self, path: str, create: bool = False):
        node = self._file if path.startswith("/") else self
        parts = [p for p in path.split("/") if p]
        for i, p in enumerate(parts):
            if not isinstance(node, Group):
                raise KeyError(path)
            if p not in node._children:
                if not create:
                    raise KeyError(f"Unable to open object (object '{p}' doesn't exist)")
                node._file._require_writable()
                node._children[p] = Group(node._file, (node.name.rstrip("/") + "/" + p))
                node._file._dirty = True
            node = node._children[p]
        return node

    def __getitem__(self, path: str):
        return self._walk(path)

    def __contains__(self, path) -> bool:
        try:
            self._walk(str(path))
            return True
        except KeyError:
            return False

    def __delitem__(self, path: str) -> None:
        self._file._require_writable()
        parent, _, leaf = path.rstrip("/").rpartition("/")
        grp = self._walk(parent) if parent else self
        del grp._children[leaf]
        self._file._dirty = True

    def __iter__(self) -> Iterator[str]:
        return iter(sorted(self._children))

    def __len__(self):
        return len(self._children)

    def keys(self):
        return sorted(self._children)

    def items(self):
        return [(k, self._children[k]) for k in sorted(self._children)]

    def values(self):
        return [self._children[k] for k in sorted(self._children)]

    def get(self, path, default=None):
        try:
            return self._walk(path)
        except KeyError:
            return default

    # ---- creation ----
    def create_group(self, path: str) -> "Group":
        if path in self:
            raise ValueError(f"Unable to create group (name already exists): {path}")
        return self._walk(path, create=True)

    def require_group(self, path: str) -> "Group":
        g = self._walk(path, create=True)
        if not isinstance(g, Group):
            raise TypeError(f"Incompatible object ({type(g).__name__}) already exists")
        return g

    def create_dataset(self, name: str, shape=None, dtype=None, data=None, maxshape=None, chunks=None, **unsupported) -> Dataset:
        self._file._require_writable()
        for k in unsupported:
            if unsupported[k] not in (None, False):
                raise TypeError(f"h5lite: create_dataset option '{k}' is not supported")
        if data is not None:
            arr = np.array(data, dtype=dtype) if dtype is not None else np.array(data)
            if shape is not None and tuple(np.atleast_1d(shape)) != arr.shape:
                arr = arr.reshape(shape)
        else:
            if shape is None:
                raise TypeError("h5lite: one of shape or data must be given")
            shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
            arr = np.zeros(shape, dtype=np.dtype(dtype if dtype is not None else np.float32))
        if arr.dtype.kind not in "iufS":
            raise TypeError(f"h5lite: dataset dtype {arr.dtype} is not supported (int / float / fixed bytes)")
        if maxshape is None:
            maxshape = arr.shape
        maxshape = tuple(None if m is None else int(m) for m in (maxshape if isinstance(maxshape, (tuple, list)) else (maxshape,)))
        if chunks is True or (chunks is None and any(m is None or m != s for m, s in zip(maxshape, arr.shape))):
            chunks = tuple(max(1, min(s if s else 1, 1024)) for s in arr.shape)   # h5py guesses; any valid chunking will do
        if chunks is not None:
            chunks = tuple(int(c) for c in (chunks if isinstance(chunks, (tuple, list)) else (chunks,)))
            if len(chunks) != arr.ndim or any(c <= 0 for c in chunks):
                raise ValueError("h5lite: chunks must match the dataset rank and be positive")
        parent, _, leaf = name.rstrip("/").rpartition("/")
        grp = self._walk(parent, create=True) if parent else self
        if leaf in grp._children:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        ds = Dataset(self._file, grp.name.rstrip("/") + "/" + leaf, arr, maxshape, chunks)
        grp._children[leaf] = ds
        self._file._dirty = True
        return ds

    def move(self, src: str, dst: str) -> None:
        self._file._require_writable()
        if dst in self:
            raise ValueError(f"Unable to move link (destination '{dst}' exists)")
        sp, _, sl = src.rstrip("/").rpartition("/")
        dp, _, dl = dst.rstrip("/").rpartition("/")
        sg = self._walk(sp) if sp else self
        dg = self._walk(dp, create=True) if dp else self
        node = sg._children.pop(sl)
        node.name = dg.name.rstrip("/") + "/" + dl
        dg._children[dl] = node
        self._file._dirty = True


class File(Group):
    """h5py.File look-alike.  modes: "r", "r+", "w", "w-" / "x", "a"."""

    def __init__(self, path, mode: str = "r", **_ignored):
        self._path = os.fspath(path)
        self._mode = mode
        self._dirty = False
        self._open = True
        super().__init__(self, "/")
        self._file = self
        exists = os.path.exists(self._path)
        if mode in ("r", "r+"):
            if not exists:
                raise FileNotFoundError(f"Unable to open file (unable to open file: name = '{self._path}')")
            _Reader(self._path).load_into(self)
        elif mode in ("w-", "x"):
            if exists:
                raise FileExistsError(f"Unable to create file (file exists): {self._path}")
            self._dirty = True
        elif mode == "w":
            self._dirty = True
        elif mode == "a":
            if exists:
                _Reader(self._path).load_into(self)
            else:
                self._dirty = True
        else:
            raise ValueError(f"h5lite: invalid mode {mode!r}")
        if mode in ("w", "w-", "x") or (mode == "a" and not exists):
            self.flush()   # h5py creates the file at open time

    @property
    def filename(self):
        return self._path

    @property
    def mode(self):
        return "r" if self._mode == "r" else "r+"

    def _require_writable(self):
        if not self._open:
            raise ValueError("h5lite: file is closed")
        if self._mode == "r":
            raise OSError("h5lite: file is open read-only")

    def flush(self) -> None:
        if self._open and self._mode != "r" and self._dirty:
            _Writer(self).write(self._path)
            self._dirty = False

    def close(self) -> None:
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __bool__(self):
        return self._open


def string_dtype(encoding: str = "utf-8", length: int | None = None):
    if length is None:
        raise TypeError("h5lite: variable-length string datasets are not supported (attributes only)")
    return np.dtype(f"S{int(length)}")


# =====================================================================================================================
# writer
# =====================================================================================================================
def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _dtype_message(dt: np.dtype) -> bytes:
    """Datatype message body, version 1."""
    dt = np.dtype(dt)
    if dt.kind in "iu":
        flags0 = 0x08 if dt.kind == "i" else 0x00               # little endian, no padding, signed bit
        return struct.pack("<BBBBI", 0x10 | 0, flags0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "f":
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            sign = 31
        elif dt.itemsize == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            sign = 63
        elif dt.itemsize == 2:
            props = struct.pack("<HHBBBBI", 0, 16, 10, 5, 0, 10, 15)
            sign = 15
        else:
            raise TypeError(f"h5lite: float{dt.itemsize * 8} unsupported")
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize) + props   # mantissa normalisation: implied msb
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, max(1, dt.itemsize))        # NULL-padded, ASCII
    raise TypeError(f"h5lite: dtype {dt} unsupported")


_VLEN_STR_DTYPE = (struct.pack("<BBBBI", 0x10 | 9, 0x01, 0x01, 0, 16)                 # vlen, type = string, null-terminated, UTF-8
                   + struct.pack("<BBBBI", 0x10 | 3, 0x00, 0, 0, 1))                   # base type: 1-byte character


def _dataspace_message(shape, maxshape=None) -> bytes:
    rank = len(shape)
    flags = 1 if maxshape is not None else 0
    b = struct.pack("<BBB5x", 1, rank, flags)
    b += b"".join(struct.pack("<Q", int(s)) for s in shape)
    if maxshape is not None:
        b += b"".join(struct.pack("<Q", UNDEF if m is None else int(m)) for m in maxshape)
    return b


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


class _Writer:
    def __init__(self, file: File):
        self.file = file
        self.buf = None
        self.pos = 0
        self.gheap: dict[int, tuple[int, int]] = {}   # id(attr owner, key) -> (collection address, index)

    # ---- raw output ----
    def _alloc(self, n: int) -> int:
        self.pos += -self.pos % 8
        addr = self.pos
        self.pos += n
        return addr

    def _put(self, addr: int, data: bytes) -> None:
        self.out.seek(addr)
        self.out.write(data)

    def _emit(self, data: bytes) -> int:
        addr = self._alloc(len(data))
        self._put(addr, data)
        return addr

    # ---- global heap: every variable-length string attribute value of the file ----
    def _collect_strings(self, node, acc):
        for k, v in node.attrs.items():
            if isinstance(v, str):
                acc.append(((id(node), k), v.encode("utf-8")))
        if isinstance(node, Group):
            for c in node._children.values():
                self._collect_strings(c, acc)

    def _write_global_heaps(self, strings):
        i = 0
        while i < len(strings):
            objs, size = [], 16                      # collection header
            while i < len(strings) and len(objs) < 0xFFFE:
                need = 16 + len(_pad8(strings[i][1]))
                if objs and size + need + 16 > 65536:
                    break
                objs.append(strings[i])
                size += need
                i += 1
            total = max(4096, size + 16)             # + the trailing free-space object; H5HG_MINSIZE = 4096
            total += -total % 8
            addr = self._alloc(total)
            b = bytearray(b"GCOL" + struct.pack("<B3xQ", 1, total))
            for idx, (key, data) in enumerate(objs, start=1):
                b += struct.pack("<HH4xQ", idx, 1, len(data)) + _pad8(data)   # refcount 1; index 0 is reserved for free space
                self.gheap[key] = (addr, idx)
            free = total - len(b)
            b += struct.pack("<HH4xQ", 0, 0, free)   # object 0: the remaining free space (size includes this header)
            b += b"\x00" * (total - len(b))
            self._put(addr, bytes(b))

    # ---- attributes ----
    def _attr_message(self, owner, name: str, v: Any) -> bytes:
        nm = name.encode("utf-8") + b"\x00"
        if isinstance(v, str):
            dt, ds = _VLEN_STR_DTYPE, _dataspace_message(())
            caddr, idx = self.gheap[(id(owner), name)]
            data = struct.pack("<IQI", len(v.encode("utf-8")), caddr, idx)
        else:
            a = np.asarray(v)
            if a.dtype.kind == "S" and a.dtype.itemsize == 0:
                a = a.astype("S1")
            a = a.astype(a.dtype.newbyteorder("<")) if a.dtype.kind in "iuf" else a
            dt, ds = _dtype_message(a.dtype), _dataspace_message(a.shape)
            data = a.tobytes()
        body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data
        return _message(0x000C, body)

    def _object_header(self, messages: list[bytes]) -> int:
        body = b"".join(messages)
        hdr = struct.pack("<BxHII4x", 1, len(messages), 1, len(body))   # version 1, #messages, refcount 1, header size; pad to 8
        return self._emit(hdr + body)

    # ---- datasets ----
    def _write_chunk_btree(self, entries, rank: int, chunk_bytes: int, end_key) -> int:
        """entries: [(offsets tuple, address)] sorted; returns the root node address (B-tree v1, node type 1)."""
        key_size = 8 + 8 * (rank + 1)
        cap = 2 * ISTORE_K
        node_size = 24 + (cap + 1) * key_size + cap * 8

        def key(offs, nbytes):
            return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", int(o)) for o in offs) + struct.pack("<Q", 0)

        level = 0
        items = [(offs, addr) for offs, addr in entries]            # (first key offsets, child address)
        while True:
            groups = [items[i:i + cap] for i in range(0, len(items), cap)]
            addrs = [self._alloc(node_size) for _ in groups]
            for gi, grp in enumerate(groups):
                b = bytearray(b"TREE" + struct.pack("<BBH", 1, level, len(grp)))
                b += struct.pack("<QQ", addrs[gi - 1] if gi > 0 else UNDEF, addrs[gi + 1] if gi + 1 < len(groups) else UNDEF)
                for offs, child in grp:
                    b += key(offs, chunk_bytes) + struct.pack("<Q", child)
                nxt = groups[gi + 1][0][0] if gi + 1 < len(groups) else end_key
                b += key(nxt, chunk_bytes if gi + 1 < len(groups) else 0)
                b += b"\x00" * (node_size - len(b))
                self._put(addrs[gi], bytes(b))
            items = [(grp[0][0], addrs[gi]) for gi, grp in enumerate(groups)]
            if len(items) == 1:
                return items[0][1]
            level += 1

    def _write_dataset(self, ds: Dataset) -> int:
        a = ds._data
        a = a.astype(a.dtype.newbyteorder("<")) if a.dtype.kind in "iuf" else a
        a = np.ascontiguousarray(a)
        rank = a.ndim
        msgs = [_message(0x0001, _dataspace_message(a.shape, ds.maxshape)), _message(0x0003, _dtype_message(a.dtype), flags=1)]
        msgs.append(_message(0x0005, struct.pack("<BBBBI", 2, 3 if ds.chunks else 2, 2, 1, 0)))   # fill value v2: default, alloc incr/late
        if ds.chunks is not None:
            chunks = ds.chunks
            chunk_bytes = int(np.prod(chunks)) * a.dtype.itemsize
            grid = [(-(-a.shape[d] // chunks[d])) for d in range(rank)]
            entries = []
            if a.size:
                for idx in np.ndindex(*grid):
                    offs = tuple(i * c for i, c in zip(idx, chunks))
                    blk = np.zeros(chunks, dtype=a.dtype)
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, a.shape))
                    blk[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
                    entries.append((offs, self._emit(blk.tobytes())))
            if entries:
                end_key = tuple([grid[0] * chunks[0]] + [0] * (rank - 1)) if rank else ()
                baddr = self._write_chunk_btree(entries, rank, chunk_bytes, end_key)
            else:
                baddr = UNDEF
            layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", baddr)
            layout += b"".join(struct.pack("<I", int(c)) for c in chunks) + struct.pack("<I", a.dtype.itemsize)
        else:
            data = a.tobytes()
            daddr = self._emit(data) if data else UNDEF
            layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", daddr, len(data))
        msgs.append(_message(0x0008, layout))
        msgs += [self._attr_message(ds, k, v) for k, v in ds.attrs.items()]
        return self._object_header(msgs)

    # ---- groups ----
    def _write_group(self, grp: Group) -> tuple[int, int, int]:
        """returns (object header address, B-tree address, local heap address)"""
        names = sorted(grp._children, key=lambda s: s.encode("utf-8"))
        child_addr = {}
        child_stab = {}
        for n in names:
            c = grp._children[n]
            if isinstance(c, Group):
                child_addr[n], bt, hp = self._write_group(c)
                child_stab[n] = (bt, hp)
            else:
                child_addr[n] = self._write_dataset(c)
        # local heap: offset 0 holds the empty string (the B-tree's left-most key)
        heap = bytearray(b"\x00" * 8)
        name_off = {}
        for n in names:
            name_off[n] = len(heap)
            heap += _pad8(n.encode("utf-8") + b"\x00")
        heap_data_addr = self._emit(bytes(heap))
        heap_addr = self._emit(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data_addr))   # free-list head 1 = none (H5HL_FREE_NULL)
        # symbol table nodes (up to 2 * leaf K entries each), then one B-tree node over them
        per = 2 * GROUP_LEAF_K
        snods = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(snods) > 2 * GROUP_INTERNAL_K:
            raise ValueError("h5lite: more than 256 links in one group are not supported")
        snod_addrs = []
        for part in snods:
            b = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(part)))
            for n in part:
                if n in child_stab:
                    b += struct.pack("<QQII", name_off[n], child_addr[n], 1, 0) + struct.pack("<QQ", *child_stab[n])
                else:
                    b += struct.pack("<QQII", name_off[n], child_addr[n], 0, 0) + b"\x00" * 16
            b += b"\x00" * (8 + per * 40 - len(b))
            snod_addrs.append(self._emit(bytes(b)))
        cap = 2 * GROUP_INTERNAL_K
        node = bytearray(b"TREE" + struct.pack("<BBH", 0, 0, len(snods) if names else 0) + struct.pack("<QQ", UNDEF, UNDEF))
        node += struct.pack("<Q", 0)
        if names:
            for part, sa in zip(snods, snod_addrs):
                node += struct.pack("<QQ", sa, name_off[part[-1]])   # child, then the key = largest name in that child
        node += b"\x00" * (24 + (cap + 1) * 8 + cap * 8 - len(node))
        btree_addr = self._emit(bytes(node))
        msgs = [_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [self._attr_message(grp, k, v) for k, v in grp.attrs.items()]
        return self._object_header(msgs), btree_addr, heap_addr

    def write(self, path: str) -> None:
        tmp = f"{path}.h5lite-{os.getpid()}"
        with open(tmp, "wb") as out:
            self.out = out
            self.pos = 96                                  # superblock v0 with 8-byte offsets
            strings = []
            self._collect_strings(self.file, strings)
            self._write_global_heaps(strings)
            root_hdr, root_bt, root_heap = self._write_group(self.file)
            eof = self.pos + (-self.pos % 8)
            sb = _SIG + struct.pack("<BBBxBBBx", 0, 0, 0, 0, 8, 8) + struct.pack("<HHI", GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
            sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
            sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_bt, root_heap)
            assert len(sb) == 96
            self._put(0, sb)
            out.truncate(eof)
        os.replace(tmp, path)


# =====================================================================================================================
# reader (written against the specification, independently of the writer's helpers)
# =====================================================================================================================
class _Reader:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.d = f.read()
        self.base = 0
        off = 0
        while True:                                   # the superblock may follow a user block: 0, 512, 1024, ...
            if self.d[off:off + 8] == _SIG:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(self.d):
                raise OSError("h5lite: not an HDF5 file (no superblock signature)")
        self.sb_off = off
        ver = self.d[off + 8]
        if ver not in (0, 1):
            raise NotImplementedError(f"h5lite: superblock version {ver} (new-style groups) is not supported; "
                                      "h5py writes version 0 unless libver='latest' is requested")
        so, sl = self.d[off + 13], self.d[off + 14]
        if (so, sl) != (8, 8):
            raise NotImplementedError(f"h5lite: only 8-byte offsets / lengths are supported (file has {so}/{sl})")
        p = off + 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", self.d, p)
        # every address in the file is relative to this base address (spec III.A), not to where the signature was found
        p += 32
        _name_off, self.root_hdr, _ctype = struct.unpack_from("<QQI", self.d, p)

    def _at(self, addr: int) -> int:
        return addr + self.base

    # ---- object headers ----
    def _messages(self, addr: int):
        p = self._at(addr)
        ver, nmsg, _ref, hsize = struct.unpack_from("<BxHII", self.d, p)
        if ver != 1:
            raise NotImplementedError(f"h5lite: object header version {ver} is not supported")
        blocks = [(p + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.d, q)
                body = self.d[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x0010:                       # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self._at(caddr), clen))
                out.append((mtype, flags, body))
        return out

    @staticmethod
    def _parse_dataspace(b):
        ver, rank, flags = b[0], b[1], b[2]
        p = 8 if ver == 1 else 4
        dims = struct.unpack_from(f"<{rank}Q", b, p)
        p += 8 * rank
        maxd = None
        if flags & 1:
            maxd = tuple(None if m == UNDEF else int(m) for m in struct.unpack_from(f"<{rank}Q", b, p))
        return tuple(int(x) for x in dims), maxd

    @staticmethod
    def _parse_dtype(b):
        """-> (numpy dtype or "vlen_str", size)"""
        cls, ver = b[0] & 0x0F, b[0] >> 4
        f0, f1 = b[1], b[2]
        size = struct.unpack_from("<I", b, 4)[0]
        order = ">" if (f0 & 1) else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if f0 & 0x08 else 'u'}{size}"), size
        if cls == 1:
            return np.dtype(f"{order}f{size}"), size
        if cls == 3:
            return np.dtype(f"S{size}"), size
        if cls == 9 and (f0 & 0x0F) == 1:
            return "vlen_str", size
        return None, size

    def _global_heap_object(self, caddr: int, index: int) -> bytes:
        p = self._at(caddr)
        if self.d[p:p + 4] != b"GCOL":
            raise OSError("h5lite: bad global heap collection signature")
        total = struct.unpack_from("<Q", self.d, p + 8)[0]
        q, end = p + 16, p + total
        while q + 16 <= end:
            idx, _ref, size = struct.unpack_from("<HH4xQ", self.d, q)
            if idx == index:
                return self.d[q + 16:q + 16 + size]
            if idx == 0:
                break
            q += 16 + size + (-size % 8)
        raise KeyError(f"h5lite: global heap object {index} not found")

    def _parse_attribute(self, b):
        ver = b[0]
        nsize, tsize, ssize = struct.unpack_from("<HHH", b, 2)
        p = 8
        pad = (lambda n: n + (-n % 8)) if ver == 1 else (lambda n: n)
        if ver == 3:
            p += 1                                           # name character set
        name = b[p:p + nsize].split(b"\x00")[0].decode("utf-8")
        p += pad(nsize)
        dt, esize = self._parse_dtype(b[p:p + tsize])
        p += pad(tsize)
        shape, _ = self._parse_dataspace(b[p:p + ssize]) if ssize else ((), None)
        p += pad(ssize)
        n = int(np.prod(shape)) if shape else 1
        raw = b[p:p + n * esize]
        if dt == "vlen_str":
            vals = []
            for i in range(n):
                ln, caddr, idx = struct.unpack_from("<IQI", raw, i * 16)
                vals.append(self._global_heap_object(caddr, idx)[:ln].decode("utf-8") if ln else "")
            return name, (vals[0] if not shape else np.asarray(vals, dtype=object).reshape(shape))
        if dt is None:
            return name, None
        arr = np.frombuffer(raw, dtype=dt, count=n)
        arr = arr.astype(arr.dtype.newbyteorder("=")) if arr.dtype.kind in "iuf" else arr
        return name, (arr[0] if not shape else arr.reshape(shape).copy())

    # ---- B-trees ----
    def _group_entries(self, btree_addr: int, heap_addr: int):
        hp = self._at(heap_addr)
        if self.d[hp:hp + 4] != b"HEAP":
            raise OSError("h5lite: bad local heap signature")
        _dsize, _free, daddr = struct.unpack_from("<QQQ", self.d, hp + 8)
        dbase = self._at(daddr)

        def name_at(off):
            e = self.d.index(b"\x00", dbase + off)
            return self.d[dbase + off:e].decode("utf-8")

        out = []

        def walk(addr):
            p = self._at(addr)
            sig = self.d[p:p + 4]
            if sig == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", self.d, p + 4)
                if ntype != 0:
                    raise OSError("h5lite: expected a group B-tree node")
                q = p + 24 + 8
                for _ in range(used):
                    child = struct.unpack_from("<Q", self.d, q)[0]
                    walk(child)
                    q += 16
            elif sig == b"SNOD":
                nsym = struct.unpack_from("<H", self.d, p + 6)[0]
                q = p + 8
                for _ in range(nsym):
                    noff, oaddr, ctype = struct.unpack_from("<QQI", self.d, q)
                    out.append((name_at(noff), oaddr))
                    q += 40
            else:
                raise OSError(f"h5lite: unexpected node signature {sig!r}")

        walk(btree_addr)
        return out

    def _chunk_entries(self, addr: int, rank: int):
        out = []
        key_size = 8 + 8 * (rank + 1)

        def walk(a):
            p = self._at(a)
            if self.d[p:p + 4] != b"TREE":
                raise OSError("h5lite: bad chunk B-tree node")
            ntype, level, used = struct.unpack_from("<BBH", self.d, p + 4)
            if ntype != 1:
                raise OSError("h5lite: expected a raw-data chunk B-tree node")
            q = p + 24
            for _ in range(used):
                nbytes, fmask = struct.unpack_from("<II", self.d, q)
                offs = struct.unpack_from(f"<{rank}Q", self.d, q + 8)
                child = struct.unpack_from("<Q", self.d, q + key_size)[0]
                if level == 0:
                    out.append((offs, child, nbytes, fmask))
                else:
                    walk(child)
                q += key_size + 8

        walk(addr)
        return out

    # ---- datasets ----
    def _read_dataset(self, file, name, msgs) -> Dataset:
        shape = maxshape = dt = None
        layout = None
        filters = []
        attrs = []
        for mtype, _flags, b in msgs:
            if mtype == 0x0001:
                shape, maxshape = self._parse_dataspace(b)
            elif mtype == 0x0003:
                dt, _ = self._parse_dtype(b)
            elif mtype == 0x0008:
                layout = b
            elif mtype == 0x000B:
                filters = self._parse_filters(b)
            elif mtype == 0x000C:
                attrs.append(self._parse_attribute(b))
        if shape is None or layout is None:
            raise OSError(f"h5lite: dataset '{name}' lacks a dataspace or layout message")
        if dt is None or isinstance(dt, str):
            raise NotImplementedError(f"h5lite: dataset '{name}' has an unsupported datatype")
        n = int(np.prod(shape)) if shape else 1
        chunks = None
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 2:
                ndim = layout[2]
                baddr = struct.unpack_from("<Q", layout, 3)[0]
                cd = struct.unpack_from(f"<{ndim}I", layout, 11)
                chunks = tuple(int(c) for c in cd[:-1])
                arr = np.zeros(shape, dtype=dt)
                if baddr != UNDEF and n:
                    for offs, caddr, nbytes, fmask in self._chunk_entries(baddr, len(shape)):
                        raw = self.d[self._at(caddr):self._at(caddr) + nbytes]
                        raw = self._unfilter(raw, filters, fmask, dt.itemsize)
                        blk = np.frombuffer(raw, dtype=dt, count=int(np.prod(chunks))).reshape(chunks)
                        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, shape))
                        if any(s.stop <= s.start for s in sl):
                            continue
                        arr[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
            elif cls == 1:
                daddr, dsize = struct.unpack_from("<QQ", layout, 2)
                arr = (np.frombuffer(self.d, dtype=dt, count=n, offset=self._at(daddr)).reshape(shape).copy()
                       if daddr != UNDEF and n else np.zeros(shape, dtype=dt))
            elif cls == 0:
                dsize = struct.unpack_from("<H", layout, 2)[0]
                arr = np.frombuffer(layout[4:4 + dsize], dtype=dt, count=n).reshape(shape).copy()
            else:
                raise NotImplementedError(f"h5lite: layout class {cls}")
        elif ver in (1, 2):
            ndim, cls = layout[1], layout[2]
            p = 8
            daddr = UNDEF
            if cls != 0:
                daddr = struct.unpack_from("<Q", layout, p)[0]
                p += 8
            dims = struct.unpack_from(f"<{ndim}I", layout, p)
            if cls == 1:
                arr = (np.frombuffer(self.d, dtype=dt, count=n, offset=self._at(daddr)).reshape(shape).copy()
                       if daddr != UNDEF and n else np.zeros(shape, dtype=dt))
            else:
                raise NotImplementedError(f"h5lite: layout v{ver} class {cls}")
            del dims
        else:
            raise NotImplementedError(f"h5lite: data layout version {ver}")
        if arr.dtype.kind in "iuf":
            arr = arr.astype(arr.dtype.newbyteorder("="))
        ds = Dataset(file, name, arr, maxshape if maxshape is not None else shape, chunks)
        for k, v in attrs:
            ds.attrs._d[k] = v
        return ds

    @staticmethod
    def _parse_filters(b):
        ver, nf = b[0], b[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid, nlen, _flags, ncd = struct.unpack_from("<HHHH", b, p)
            p += 8
            if ver == 1 or fid >= 256:
                p += nlen + (-nlen % 8 if ver == 1 else 0)
            cd = struct.unpack_from(f"<{ncd}I", b, p)
            p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
            out.append((fid, cd))
        return out

    @staticmethod
    def _unfilter(raw, filters, fmask, itemsize):
        for i in reversed(range(len(filters))):
            if fmask & (1 << i):
                continue
            fid, _cd = filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                a = np.frombuffer(raw, dtype=np.uint8)
                nel = len(a) // itemsize
                raw = a[:nel * itemsize].reshape(itemsize, nel).T.tobytes() + a[nel * itemsize:].tobytes()
            elif fid == 3:
                raw = raw[:-4]                                # fletcher32 checksum
            else:
                raise NotImplementedError(f"h5lite: filter {fid} is not supported")
        return raw

    # ---- tree ----
    def _load_object(self, file, parent: Group, leaf: str, addr: int):
        msgs = self._messages(addr)
        stab = [b for t, _f, b in msgs if t == 0x0011]
        path = parent.name.rstrip("/") + "/" + leaf if parent is not None else "/"
        if stab:
            grp = file if parent is None else Group(file, path)
            bt, hp = struct.unpack_from("<QQ", stab[0], 0)
            for t, _f, b in msgs:
                if t == 0x000C:
                    k, v = self._parse_attribute(b)
                    grp.attrs._d[k] = v
            for name, oaddr in self._group_entries(bt, hp):
                self._load_object(file, grp, name, oaddr)
            if parent is not None:
                parent._children[leaf] = grp
        elif any(t == 0x0008 for t, _f, _b in msgs):
            try:
                parent._children[leaf] = self._read_dataset(file, path, msgs)
            except NotImplementedError:
                pass                                       # leave unsupported objects out of the tree
        elif any(t in (0x0002, 0x0006) for t, _f, _b in msgs):
            raise NotImplementedError("h5lite: new-style groups (link messages) are not supported")

    def load_into(self, file: File) -> None:
        self._load_object(file, None, "", self.root_hdr)


def is_hdf5(path) -> bool:
    try:
        with open(path, "rb") as f:
            return f.read(8) == _SIG
    except OSError:
        return False
