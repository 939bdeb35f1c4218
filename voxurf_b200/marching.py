"""GPU marching cubes for the mesh export: the step after the field query of
lib/voxurf_fine.py:894-910 -- the reference hands the res^3 field to PyMCubes (`mcubes.marching_cubes(u, threshold)`,
lib/dvgo_ori.py:695-703), a host library that dominates end-to-end mesh time at 512^3.  PyMCubes is not vendored in the
reference and not installed here, so its exact triangulation cannot be pinned ("parity unpinned" for triangle order /
ambiguous-face choices); what is reproduced is its contract: indexed mesh, one vertex per lattice edge crossed by the
iso-level, linear interpolation along the edge, vertices in lattice-index coordinates (extract_geometry then maps them
with `vertices / (resolution - 1) * (bound_max - bound_min) + bound_min`).

The 256-case triangle table is GENERATED here (no copied table): on every cube face the crossings of the face's four
edges are joined into segments -- two crossings: one segment; four crossings (ambiguous face): the two segments that cut
off the face's two above-threshold corners -- a rule that depends on the face's corner signs only, so the two cubes that
share a face agree and the surface is watertight.  Segments are chained into closed polygons and fanned into triangles,
oriented so that normals point from above-threshold to below-threshold corners.
"""
import numpy as np
import torch

from ._lib import call

# cube corners: bit 0 = +x (slowest lattice axis i), bit 1 = +y (j), bit 2 = +z (k, fastest)
_CORNER = [(c & 1, (c >> 1) & 1, (c >> 2) & 1) for c in range(8)]
# the 12 edges as (corner a, corner b): 4 along x, 4 along y, 4 along z
_EDGES = [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (1, 3), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7)]
_EDGE_ID = {frozenset(e): i for i, e in enumerate(_EDGES)}
# the 6 faces as corner cycles (walked so that consecutive corners share an edge)
_FACES = [(0, 2, 6, 4), (1, 5, 7, 3), (0, 4, 5, 1), (2, 3, 7, 6), (0, 1, 3, 2), (4, 6, 7, 5)]


def _face_segments(inside, face):
    """DIRECTED segments (edge id -> edge id) of the iso-contour on one face, the above-threshold region of the face on the
    LEFT of the direction of travel as seen from outside the cube; `inside[c]`: corner c is above the threshold"""
    cyc = list(face)
    pos = np.array([_CORNER[c] for c in cyc], float)
    n_f = pos.mean(0) - 0.5                                        # outward normal of the face (cube centre = 0.5)
    mid = lambda k: (pos[k] + pos[(k + 1) % 4]) / 2.0
    eid = lambda k: _EDGE_ID[frozenset((cyc[k], cyc[(k + 1) % 4]))]
    cross = [k for k in range(4) if inside[cyc[k]] != inside[cyc[(k + 1) % 4]]]     # face edge k joins cyc[k], cyc[k+1]

    def directed(ka, kb, c_in):
        a, b = mid(ka), mid(kb)
        left = np.cross(n_f, b - a)
        return (eid(ka), eid(kb)) if np.dot(left, pos[c_in] - a) > 0 else (eid(kb), eid(ka))
    if len(cross) == 2:
        k_in = [k for k in range(4) if inside[cyc[k]]]
        # the inside corner next to the first crossing
        c_in = cross[0] if inside[cyc[cross[0]]] else (cross[0] + 1) % 4
        assert c_in in k_in
        return [directed(cross[0], cross[1], c_in)]
    if len(cross) == 4:       # ambiguous face: cut off each above-threshold corner with its own segment
        return [directed((k - 1) % 4, k, k) for k in range(4) if inside[cyc[k]]]
    return []


def build_tables():
    """-> tri_table (256, 16) int32 (edge ids, -1 terminated; at most 5 triangles), tri_count (256,) int32"""
    table = -np.ones((256, 16), np.int32)
    count = np.zeros(256, np.int32)
    for case in range(256):
        inside = [(case >> c) & 1 == 1 for c in range(8)]
        nxt = {}
        for f in _FACES:
            for a, b in _face_segments(inside, f):
                assert a not in nxt, case
                nxt[a] = b
        tris, seen = [], set()
        for start in sorted(nxt):
            if start in seen:
                continue
            loop, cur = [], start
            while cur not in seen:
                seen.add(cur)
                loop.append(cur)
                cur = nxt[cur]
            assert cur == start and len(loop) >= 3, case
            # the loop runs counter-clockwise around the above-threshold region seen from outside the cube: fanning it in
            # REVERSE order makes the normals point away from that region (from u > threshold to u < threshold)
            loop = loop[::-1]
            for t in range(1, len(loop) - 1):
                tris.append([loop[0], loop[t], loop[t + 1]])
        assert len(tris) <= 5, (case, len(tris))
        count[case] = len(tris)
        for i, t in enumerate(tris):
            table[case, 3 * i:3 * i + 3] = t
    return table, count


_TABLES = None


def tables(device):
    global _TABLES
    if _TABLES is None:
        _TABLES = build_tables()
    return torch.from_numpy(_TABLES[0]).to(device), torch.from_numpy(_TABLES[1]).to(device)


@torch.no_grad()
def marching_cubes(u, threshold=0.0):
    """u (nx, ny, nz) float32 CUDA -> vertices (V, 3) float32 in lattice-index coordinates, triangles (T, 3) int64, like
    mcubes.marching_cubes(u, threshold) (lib/dvgo_ori.py:697): a corner is "inside" where u > threshold (u = -sdf)."""
    assert u.is_cuda and u.dim() == 3 and u.dtype == torch.float32
    u = u.contiguous()
    nx, ny, nz = u.shape
    dev = u.device
    tri_table, tri_count = tables(dev)
    n_cells = (nx - 1) * (ny - 1) * (nz - 1)
    if n_cells <= 0:
        return torch.zeros(0, 3, device=dev), torch.zeros(0, 3, dtype=torch.int64, device=dev)
    cell_tris = torch.empty(n_cells, dtype=torch.int32, device=dev)
    edge_flag = torch.empty(3 * u.numel(), dtype=torch.uint8, device=dev)
    call('vx_mc_classify', u, nx, ny, nz, float(threshold), tri_count, cell_tris, edge_flag)
    tri_off = torch.cumsum(cell_tris, 0, dtype=torch.int64)
    vert_off = torch.cumsum(edge_flag, 0, dtype=torch.int32)          # inclusive: id of a flagged edge = value - 1
    n_tri, n_vert = int(tri_off[-1]), int(vert_off[-1])               # the one host read of the extraction
    verts = torch.empty(n_vert, 3, dtype=torch.float32, device=dev)
    tris = torch.empty(n_tri, 3, dtype=torch.int64, device=dev)
    if n_tri:
        call('vx_mc_emit', u, nx, ny, nz, float(threshold), tri_table, tri_off, vert_off, edge_flag, verts, tris)
    return verts, tris


def extract_geometry(model, resolution=128, threshold=0.0, smooth=True, sigma=0.5):
    """lib/voxurf_fine.py:894-910 + lib/dvgo_ori.py:695-703 on the GPU: field query, marching cubes, vertices mapped to world
    coordinates.  -> (vertices (V,3) float32 CUDA, triangles (T,3) int64 CUDA)"""
    u = model.query_sdf_field(resolution, smooth=smooth, sigma=sigma)
    v, t = marching_cubes(u, threshold)
    mn = torch.tensor(model._min_host, device=v.device)
    mx = torch.tensor(model._max_host, device=v.device)
    return v / (resolution - 1.0) * (mx - mn)[None] + mn[None], t
