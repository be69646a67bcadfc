"""ORACLE — TEST INFRASTRUCTURE ONLY.

Independent numpy parsers for the reference's two map file formats, used to check the product's
C++ loaders (trajectory_planner_b200/csrc/tp_map.cpp) and to build oracle maps.

* PCD (map/square_static_map.pcd): ASCII v0.7, header lines until `DATA ascii`, then `x y z`
  float rows.  Consumed by mapManager::occMap's prebuilt-map loader in the
  reference (cfg/bspline_interactive/occupancy_map.yaml:50).
* OctoMap binary tree (.bt, map/*.bt): text header (`id OcTree`, `size`, `res`, `data`), then
  a depth-first stream of 2 bytes per inner node = 8 children x 2 bits, LSB first, child index
  bit0->x bit1->y bit2->z; codes 0 unknown, 1 free leaf, 2 occupied leaf, 3 inner node;
  16 levels, key origin 32768 (octomap's OcTreeBaseImpl::readBinaryNode; call sites
  polyTrajOctomap.cpp:142-143, 573-599).
"""
import struct

import numpy as np


def read_pcd_ascii(path):
    """-> float64 array [n,3] of the float32 points in the file."""
    with open(path, "rb") as f:
        raw = f.read()
    pos = raw.index(b"DATA ascii")
    pos = raw.index(b"\n", pos) + 1
    header = raw[:pos].decode()
    npts = None
    for line in header.splitlines():
        if line.startswith("POINTS"):
            npts = int(line.split()[1])
    # the reference loads these through PCL as float32
    pts = np.array(raw[pos:].split(), dtype=np.float32).reshape(-1, 3)
    assert npts is None or npts == len(pts)
    return pts.astype(np.float64)


def read_bt(path):
    """-> (res, leaves) with leaves = int64 array [n,5]: key_x, key_y, key_z (of the leaf's min
    corner, 16-bit key space, origin 32768), size (in finest cells, power of two), occupied(0/1)."""
    with open(path, "rb") as f:
        raw = f.read()
    pos = raw.index(b"data\n") + 5
    header = raw[:pos].decode(errors="replace")
    res = None
    size = None
    for line in header.splitlines():
        if line.startswith("res"):
            res = float(line.split()[1])
        if line.startswith("size"):
            size = int(line.split()[1])
    data = np.frombuffer(raw, dtype=np.uint8, offset=pos)
    leaves = []
    cursor = 0
    n_nodes = 1  # root
    # iterative DFS: stack of (kx, ky, kz, size_of_this_node)
    stack = [(0, 0, 0, 1 << 16)]
    while stack:
        kx, ky, kz, s = stack.pop()
        b0 = int(data[cursor])
        b1 = int(data[cursor + 1])
        cursor += 2
        bits = b0 | (b1 << 8)
        h = s >> 1
        inner = []
        for c in range(8):
            code = (bits >> (2 * c)) & 3
            if code == 0:
                continue
            cx = kx + (h if (c & 1) else 0)
            cy = ky + (h if (c & 2) else 0)
            cz = kz + (h if (c & 4) else 0)
            n_nodes += 1
            if code == 3:
                inner.append((cx, cy, cz, h))
            else:
                leaves.append((cx, cy, cz, h, 1 if code == 2 else 0))
        # children are serialised in index order, depth first
        for item in reversed(inner):
            stack.append(item)
    assert cursor == len(data), (cursor, len(data))
    assert size is None or size == n_nodes, (size, n_nodes)
    return res, np.array(leaves, dtype=np.int64).reshape(-1, 5)


def bt_to_cells(leaves, max_leaf_cells=1 << 30):
    """Expand leaves to finest-resolution cells.  -> (occ_keys [n,3], free_keys [m,3]) int64."""
    occ, free = [], []
    for kx, ky, kz, s, o in leaves:
        if s == 1:
            (occ if o else free).append(np.array([[kx, ky, kz]], dtype=np.int64))
        else:
            if s ** 3 > max_leaf_cells:
                continue
            r = np.arange(s, dtype=np.int64)
            g = np.stack(np.meshgrid(kx + r, ky + r, kz + r, indexing="ij"), -1).reshape(-1, 3)
            (occ if o else free).append(g)
    cat = lambda l: np.concatenate(l, 0) if l else np.zeros((0, 3), np.int64)
    return cat(occ), cat(free)


def bt_bbox(leaves, res, occupied_only=False):
    """metric bbox of known (or occupied) leaves: (min[3], max[3])."""
    l = leaves[leaves[:, 4] == 1] if occupied_only else leaves
    lo = (l[:, :3].min(0) - 32768) * res
    hi = ((l[:, :3] + l[:, 3:4]).max(0) - 32768) * res
    return lo, hi


def read_tpm(path):
    """This repo's compact raster format (trajectory_planner_b200/csrc/tp_map.cpp: "TPM1" | pad i32 | res f64 | origin
    3 x f64 | dims 3 x i32 + pad | two RLE streams over the z-fastest bit-packed words: occupied, known).
    -> dict(res, origin[3], dims[3], occupied[nx,ny,nz] uint8, known[nx,ny,nz] uint8).  Numpy only: lets the
    reference arm of bench.py build its map without loading the product library."""
    raw = open(path, "rb").read()
    assert raw[:4] == b"TPM1", path
    pos = 8
    res = struct.unpack_from("<d", raw, pos)[0]; pos += 8
    origin = np.frombuffer(raw, "<f8", 3, pos).copy(); pos += 24
    dims = np.frombuffer(raw, "<i4", 4, pos)[:3].astype(int); pos += 16
    wz = (dims[2] + 31) // 32
    nw = int(dims[0]) * int(dims[1]) * int(wz)
    grids = []
    for _ in range(2):
        npairs = struct.unpack_from("<Q", raw, pos)[0]; pos += 8
        pairs = np.frombuffer(raw, "<u4", 2 * npairs, pos).reshape(-1, 2); pos += 8 * npairs
        words = np.repeat(pairs[:, 1], pairs[:, 0])
        assert len(words) == nw
        bits = ((words[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.uint8)
        grids.append(bits.reshape(dims[0], dims[1], wz * 32)[:, :, :dims[2]].copy())
    return dict(res=float(res), origin=origin, dims=dims, occupied=grids[0], known=grids[1])
