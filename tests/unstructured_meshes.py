"""TEST INFRASTRUCTURE: non-Cartesian (mapped Q1) quadrilateral meshes for SURVEY 8(f) row 3.

The product consumes stencil arrays through its importer (ryujin_b200/stencil_io.py); the arrays of a
mapped mesh are assembled by the oracle's quadrature assembly (oracle/orc_mesh_quads.h, a restatement
of offline_data.template.h:420-1105 for MappingQ1 quads), which stands in for a ryujin + deal.II run
that would dump them.  Geometries:

  forward_facing_step_rounded(r)   source/geometry_step.h:36-128: the Mach-3 step WITH the rounded
                                   corner (radius 0.0125): the three level-4 cells around the corner
                                   are mapped, their children follow the arc (transfinite interpolation)
  cylinder_o_grid(n_r, n_theta)    O-grid around a cylinder (every cell non-affine), cf. the disc part
                                   of source/geometry_cylinder.h

DoFs are renumbered with reverse Cuthill-McKee so that neighbours of consecutive rows are close
(the reference renumbers with Cuthill-McKee too, offline_data.template.h:210-249).
"""
import numpy as np

SLIP, DIRICHLET, DO_NOTHING = 2, 4, 0


def structured_cells(nx, ny, mask=None):
    """(vertex index grid, cells) of an nx x ny cell grid, vertices lexicographic (x fastest);
    cells in deal.II's local order v0 (0,0), v1 (1,0), v2 (0,1), v3 (1,1)"""
    vid = np.arange((nx + 1) * (ny + 1)).reshape(ny + 1, nx + 1)
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    cells = np.stack([vid[jj, ii], vid[jj, ii + 1], vid[jj + 1, ii], vid[jj + 1, ii + 1]], axis=-1).reshape(-1, 4)
    if mask is not None:
        cells = cells[np.asarray(mask, dtype=bool).ravel()]
    return vid, cells


def compress_vertices(vertices, cells):
    """drop vertices no cell uses; -> (vertices, cells)"""
    used = np.zeros(len(vertices), dtype=bool)
    used[cells.ravel()] = True
    new = np.cumsum(used) - 1
    return vertices[used], new[cells]


def boundary_faces(cells):
    """boundary faces in the order deal.II visits them -- cells ascending, faces x-, x+, y-, y+ of a
    cell, the two vertices of a face in ascending local order -- because the order of a DoF's
    boundary-map entries is the order in which its boundary conditions are applied
    (offline_data.template.h:1173-1364, hyperbolic_module.template.h:122-144).
    -> (faces (f, 2), flip (f,)): the outward normal of a -> b is (dy, -dx) / |ab|, reversed if flip"""
    local = ((0, 2, True), (1, 3, False), (0, 1, False), (2, 3, True))  # deal.II faces 0..3 of a quad
    n = int(cells.max()) + 1
    all_keys = []
    for lo, hi, _ in local:
        a, b = cells[:, lo].astype(np.int64), cells[:, hi].astype(np.int64)
        all_keys.append(np.minimum(a, b) * n + np.maximum(a, b))
    keys = np.stack(all_keys, axis=1)  # (cells, 4) undirected edge keys
    uniq, counts = np.unique(keys.ravel(), return_counts=True)
    on_boundary = counts[np.searchsorted(uniq, keys)] == 1
    cell_idx, face_no = np.nonzero(on_boundary)  # row-major: cells ascending, faces ascending
    lo = np.array([f[0] for f in local])[face_no]
    hi = np.array([f[1] for f in local])[face_no]
    fl = np.array([f[2] for f in local])[face_no]
    faces = np.stack([cells[cell_idx, lo], cells[cell_idx, hi]], axis=1)
    return faces, fl.astype(np.uint8)


def rcm_permutation(cells, n):
    """new index of every vertex under reverse Cuthill-McKee of the vertex graph"""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee

    rows = np.repeat(cells, 4, axis=1).ravel()
    cols = np.tile(cells, (1, 4)).ravel()
    graph = sp.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(n, n))
    order = reverse_cuthill_mckee(graph, symmetric_mode=True)
    new_of_old = np.empty(n, dtype=np.int64)
    new_of_old[order] = np.arange(n)
    return new_of_old


def renumber(vertices, cells, new_of_old):
    out = np.empty_like(vertices)
    out[new_of_old] = vertices
    return out, new_of_old[cells]


def transfinite(bottom, top, left, right, n):
    """(n+1) x (n+1) points of the cell bounded by the four curves (callables [0,1] -> R^2, consistent
    corners): X(s, t) = (1-t) bottom(s) + t top(s) + (1-s) left(t) + s right(t) - bilinear corners"""
    s = np.linspace(0.0, 1.0, n + 1)
    B, T = np.array([bottom(x) for x in s]), np.array([top(x) for x in s])
    L, R = np.array([left(x) for x in s]), np.array([right(x) for x in s])
    out = np.zeros((n + 1, n + 1, 2))
    for j, t in enumerate(s):
        for i, u in enumerate(s):
            out[j, i] = ((1 - t) * B[i] + t * T[i] + (1 - u) * L[j] + u * R[j]
                         - ((1 - u) * (1 - t) * B[0] + u * (1 - t) * B[-1] + (1 - u) * t * T[0] + u * t * T[-1]))
    return out


def forward_facing_step_rounded(refinement, length=3.0, height=1.0, step_position=0.6, step_height=0.2, radius=0.0125):
    """-> (vertices, cells, faces, face_ids).  `refinement` counts global refinements beyond the 4 of
    geometry_step.h (0: 240 x 80 cells minus the step)."""
    coarse = 16 * 5  # level-4 cells per unit length: h4 = 0.0125 = the rounding radius
    sub = 2**refinement
    nx, ny = int(round(length * coarse)) * sub, int(round(height * coarse)) * sub
    sx, sy = int(round(step_position * coarse)) * sub, int(round(step_height * coarse)) * sub
    mask = np.ones((ny, nx), dtype=bool)
    mask[:sy, sx:] = False
    vid, cells = structured_cells(nx, ny, mask)
    X, Y = np.meshgrid(np.linspace(0.0, length, nx + 1), np.linspace(0.0, height, ny + 1))
    V = np.stack([X, Y], axis=-1)  # [j, i]

    # level-4 vertex displacements of the reference: the corner moves onto the circle around
    # (step_position + r, step_height - r), the interior neighbours above and left of it follow
    d = radius * (1.0 - np.sqrt(0.5))
    centre = np.array([step_position + radius, step_height - radius])
    h4 = 1.0 / coarse
    corner = np.array([step_position + d, step_height - d])
    above = np.array([step_position + d, step_height + h4])
    left = np.array([step_position - h4, step_height - d])

    def arc(p, q):
        """the circular arc from p to q around `centre` (both on the circle), parametrised on [0, 1]"""
        a0, a1 = np.arctan2(*(p - centre)[::-1]), np.arctan2(*(q - centre)[::-1])
        if a1 - a0 > np.pi:
            a1 -= 2 * np.pi
        if a0 - a1 > np.pi:
            a1 += 2 * np.pi
        return lambda s: centre + radius * np.array([np.cos(a0 + s * (a1 - a0)), np.sin(a0 + s * (a1 - a0))])

    def seg(p, q):
        p, q = np.asarray(p, dtype=float), np.asarray(q, dtype=float)
        return lambda s: (1 - s) * p + s * q

    wall_below = np.array([step_position, step_height - h4])  # on the vertical wall, on the circle
    wall_right = np.array([step_position + h4, step_height])  # on the top wall of the step, on the circle
    xl, xr, yb, yt = step_position - h4, step_position + h4, step_height - h4, step_height + h4
    # the three level-4 cells around the corner: (bottom, top, left, right) curves
    patches = {
        # left of the wall, below the corner height: right edge = arc from the wall point up to the corner
        (sx - sub, sy - sub): (seg([xl, yb], wall_below), seg(left, corner), seg([xl, yb], left), arc(wall_below, corner)),
        # above, left of the corner
        (sx - sub, sy): (seg(left, corner), seg([xl, yt], above), seg(left, [xl, yt]), seg(corner, above)),
        # above the step's top wall: bottom edge = arc from the corner to the wall point
        (sx, sy): (arc(corner, wall_right), seg(above, [xr, yt]), seg(corner, above), seg(wall_right, [xr, yt])),
    }
    for (i0, j0), (b, t, l, r) in patches.items():
        V[j0:j0 + sub + 1, i0:i0 + sub + 1] = transfinite(b, t, l, r, sub)
    vertices, cells = compress_vertices(V.reshape(-1, 2), cells)
    faces, flip = boundary_faces(cells)
    mid = 0.5 * (vertices[faces[:, 0]] + vertices[faces[:, 1]])
    ids = np.full(len(faces), SLIP, dtype=np.int32)
    ids[mid[:, 0] < 1e-6] = DIRICHLET
    ids[mid[:, 0] > length - 1e-6] = DO_NOTHING
    return vertices, cells, (faces, flip), ids


def cylinder_o_grid(n_r, n_theta, r_inner=0.5, r_outer=4.0, grading=1.0):
    """annulus r_inner <= r <= r_outer around the origin: slip on the cylinder, dirichlet outside"""
    t = np.linspace(0.0, 1.0, n_r + 1)
    r = r_inner * (r_outer / r_inner) ** (t**grading)  # geometric grading: nearly square cells
    th = 2.0 * np.pi * np.arange(n_theta) / n_theta
    # local (s, t) = (-theta, r) keeps the Jacobian positive: s runs clockwise
    vertices = np.stack([np.outer(r, np.cos(-th)).ravel(), np.outer(r, np.sin(-th)).ravel()], axis=1)
    vid = np.arange((n_r + 1) * n_theta).reshape(n_r + 1, n_theta)
    jj, ii = np.meshgrid(np.arange(n_r), np.arange(n_theta), indexing="ij")
    ip = (ii + 1) % n_theta
    cells = np.stack([vid[jj, ii], vid[jj, ip], vid[jj + 1, ii], vid[jj + 1, ip]], axis=-1).reshape(-1, 4)
    faces, flip = boundary_faces(cells)
    rad = np.linalg.norm(0.5 * (vertices[faces[:, 0]] + vertices[faces[:, 1]]), axis=1)
    ids = np.where(rad < 0.5 * (r_inner + r_outer), SLIP, DIRICHLET).astype(np.int32)
    return vertices, cells, (faces, flip), ids


def assemble(oracle, vertices, cells, faces, ids, reorder=True):
    """-> (oracle Offline, arrays dict, vertices in DoF order)"""
    faces, flip = faces
    if reorder:
        new_of_old = rcm_permutation(cells, len(vertices))
        vertices, cells = renumber(vertices, cells, new_of_old)
        faces = new_of_old[faces]
    off = oracle.assemble_quads(vertices, cells, faces, flip, ids)
    return off, off.arrays(), vertices


def stencil_data(arrays, vertices):
    """the oracle's arrays as a ryujin_b200.StencilData (what a deal.II run would dump)"""
    import ryujin_b200 as rj

    a = arrays
    return rj.StencilData.from_arrays(
        2, a["row_ptr"].astype(np.uint64), a["cols"], a["mij"], a["cij"], a["mi"], float(a["measure"]),
        boundary_dof=a["b_index"], boundary_normal=a["b_normal"], boundary_id=a["b_id"], boundary_position=a["b_position"],
        coupling_pairs=a["pairs"], positions=vertices)
