"""numpy restatement of the reference's VolumeOctree construction -- TEST INFRASTRUCTURE (only tests/ import it).

Follows V/src/core/datastructures/octree/volumeoctree.cpp: buildOctreeIteratively (:637-961), createTreeNodeFromTexture /
extractBrickFromTexture (:1186-1340), createParentNodeConstChannels (:1382-1536), the half-sample functors (:188-240), the uint16
conversion (:158-185) and the node statistics of volumeoctreebase.cpp:102-145 (getDepth, getNodeCount, getNumBricks).  It builds a
real pointer tree, node by node, the way the reference does; the tests then compare the product's dense per-level arrays with it.

Pinned against the reference's own known answers (V/apps/tests/octreetest/octreetest.cpp:996-2112): level counts, actual tree depth,
node counts of complete and optimised trees, the level textures against a mip map, min / max values -- see tests/test_octree.py.
One channel, uint8 / uint16 input.
"""
from __future__ import annotations

import numpy as np

NO_BRICK = None


class Node:
    __slots__ = ("avg", "mn", "mx", "brick", "children", "in_volume")

    def __init__(self, avg=0, mn=0, mx=0, brick=NO_BRICK, children=None, in_volume=False):
        self.avg, self.mn, self.mx, self.brick, self.children, self.in_volume = avg, mn, mx, brick, children, in_volume

    def depth(self):
        return 1 + max([c.depth() for c in self.children], default=0) if self.children else 1

    def node_count(self):
        return 1 + (sum(c.node_count() for c in self.children) if self.children else 0)

    def num_bricks(self):
        return (1 if self.brick is not None else 0) + (sum(c.num_bricks() for c in self.children) if self.children else 0)


def to_uint16(vol):
    if vol.dtype == np.uint16:
        return vol
    if vol.dtype == np.uint8:
        return (vol.astype(np.uint16) << 8)          # convertVoxelValueToUInt16<uint8_t> (:171-173)
    raise ValueError("uint8 / uint16 volumes only")


def halfsample(v8, fn):
    if fn == "mean":
        return int(sum(int(v) for v in v8) // 8)
    return int(max(v8)) if fn == "max" else int(min(v8))


class Octree:
    def __init__(self, vol, brick_dim, homogeneity_threshold=-1.0, fn="mean"):
        vol = to_uint16(np.ascontiguousarray(vol))
        self.vol = vol
        self.dims = vol.shape[::-1]                                     # (x, y, z)
        B = self.B = brick_dim
        assert B >= 2 and B & (B - 1) == 0
        odim = 1
        while odim < max(self.dims):
            odim <<= 1
        assert B <= odim, "brick size larger than volume dimensions"
        self.octree_dim = odim
        self.num_levels = int(np.log2(odim // B)) + 1
        self.optimize = homogeneity_threshold >= 0.0
        self.thr = int(min(max(int(np.floor(homogeneity_threshold * 65535 + 0.5)), 0), 65535))      # tgt::iround + clamp (:272-273)
        self.fn = fn
        self.root = self._build()

    # -- level 0 (full resolution) --------------------------------------------------------------------------------------------------
    def _leaf(self, n):
        B, d = self.B, self.dims
        llf = [n[a] * B for a in range(3)]
        if any(llf[a] >= d[a] for a in range(3)):
            return Node()                                               # outside the volume: empty dummy node
        s = [min(llf[a] + B, d[a]) - llf[a] for a in range(3)]
        brick = np.zeros((B, B, B), np.uint16)
        sub = self.vol[llf[2]:llf[2] + s[2], llf[1]:llf[1] + s[1], llf[0]:llf[0] + s[0]]
        brick[:s[2], :s[1], :s[0]] = sub
        total = int(sub.astype(np.uint64).sum())
        avg, mn, mx = total // (s[0] * s[1] * s[2]), int(sub.min()), int(sub.max())
        homogeneous = (mx - mn) <= self.thr
        if not homogeneous or not self.optimize:
            return Node(avg, mn, mx, brick, None, True)
        return Node(avg, mn, mx, NO_BRICK, None, True)

    # -- parents ----------------------------------------------------------------------------------------------------------------------
    def _parent(self, children, urb):
        B, H = self.B, self.B // 2
        inside = [c for c in children if c.in_volume]
        if inside:
            avg = sum(c.avg for c in inside) // len(inside)
            mn, mx = min(c.mn for c in inside), max(c.mx for c in inside)
            homogeneous = (mx - mn) <= self.thr
        else:
            avg = mn = mx = 0
            homogeneous = True
        if (homogeneous and self.optimize) or not inside:
            return Node(avg, mn, mx, NO_BRICK, None, True) if inside else Node()
        brick = np.zeros((B, B, B), np.uint16)
        for cid, child in enumerate(children):
            cp = (cid & 1, (cid >> 1) & 1, cid >> 2)                     # zyx ordered child group
            off = [cp[a] * H for a in range(3)]
            if child.brick is not None:
                llf = [min(cp[a] * B, urb[a]) for a in range(3)]
                cu = [min(llf[a] + B, urb[a]) - llf[a] for a in range(3)]
                mp = [cu[a] - 1 for a in range(3)]
                cb = child.brick
                for z in range(0, B, 2):
                    zl, zh = min(z, mp[2]), min(z + 1, mp[2])
                    for y in range(0, B, 2):
                        yl, yh = min(y, mp[1]), min(y + 1, mp[1])
                        for x in range(0, B, 2):
                            xl, xh = (x, x + 1) if x < mp[0] else (mp[0], mp[0])
                            v8 = (cb[zl, yl, xl], cb[zl, yl, xh], cb[zl, yh, xl], cb[zl, yh, xh], cb[zh, yl, xl], cb[zh, yl, xh], cb[zh, yh, xl], cb[zh, yh, xh])
                            brick[off[2] + z // 2, off[1] + y // 2, off[0] + x // 2] = halfsample(v8, self.fn)
            else:
                brick[off[2]:off[2] + H, off[1]:off[1] + H, off[0]:off[0] + H] = child.avg
        return Node(avg, mn, mx, brick, list(children), True)

    def _build(self):
        B, d = self.B, self.dims
        g = [-(-d[a] // B) for a in range(3)]                             # ceil(dims / brickDim)
        grid = {}
        for z in range(g[2]):
            for y in range(g[1]):
                for x in range(g[0]):
                    grid[(x, y, z)] = self._leaf((x, y, z))
        level = 1
        if g == [1, 1, 1]:
            return grid[(0, 0, 0)]
        while max(g) > 2:
            pg = [-(-g[a] // 2) for a in range(3)]
            cdim = [-(-d[a] // (1 << (level - 1))) for a in range(3)]
            parents = {}
            for z in range(pg[2]):
                for y in range(pg[1]):
                    for x in range(pg[0]):
                        ch = []
                        for i in range(8):
                            cpos = (x * 2 + (i & 1), y * 2 + ((i >> 1) & 1), z * 2 + (i >> 2))
                            ch.append(grid.pop(cpos) if all(cpos[a] < g[a] for a in range(3)) else Node())
                        n = (x, y, z)
                        llf = [min(n[a] * B * 2, cdim[a]) for a in range(3)]
                        urb = [min(llf[a] + 2 * B, cdim[a]) - llf[a] for a in range(3)]
                        parents[n] = self._parent(ch, urb)
            grid, g = parents, pg
            level += 1
        urb = [-(-d[a] // (1 << (level - 1))) for a in range(3)]
        ch = [grid.get((i & 1, (i >> 1) & 1, i >> 2)) or Node() for i in range(8)]
        return self._parent(ch, urb)

    # -- dense per-level view (what the product stores) ---------------------------------------------------------------------------------
    def dense(self):
        """Per level k (0 = root): avg, mn, mx (uint16), flags (uint8), bricks dict {linear index: (B, B, B) uint16}."""
        L, B = self.num_levels, self.B
        # the reference's tree may be shallower than num_levels at the top?  No: the root always sits at level 0 of num_levels levels,
        # a tree cut short by the optimisation simply has no nodes below.
        out = []
        for k in range(L):
            G = 1 << k
            out.append({"avg": np.zeros((G, G, G), np.uint16), "mn": np.zeros((G, G, G), np.uint16), "mx": np.zeros((G, G, G), np.uint16),
                        "flags": np.zeros((G, G, G), np.uint8), "bricks": {}})

        def visit(node, k, pos):
            lv = out[k]
            z, y, x = pos[2], pos[1], pos[0]
            lv["avg"][z, y, x], lv["mn"][z, y, x], lv["mx"][z, y, x] = node.avg, node.mn, node.mx
            f = 1 | (2 if node.in_volume else 0) | (4 if node.brick is not None else 0) | (8 if node.children else 0)
            lv["flags"][z, y, x] = f
            if node.brick is not None:
                lv["bricks"][(x, y, z)] = node.brick
            if node.children:
                for i, c in enumerate(node.children):
                    visit(c, k + 1, (x * 2 + (i & 1), y * 2 + ((i >> 1) & 1), z * 2 + (i >> 2)))
        visit(self.root, 0, (0, 0, 0))
        return out

    def level_volume(self, level_from_finest):
        """The volume the octree represents at a mip level (0 = full resolution): bricks where nodes have them, node averages where
        an optimised node stands in (what checkOctreeAgainstMipMap compares, octreetest.cpp:398-470)."""
        L, B = self.num_levels, self.B
        k_target = L - 1 - level_from_finest
        dims = [-(-self.dims[a] // (1 << level_from_finest)) for a in range(3)]
        vol = np.zeros(dims[::-1], np.uint16)

        def visit(node, k, pos):
            span = B << (k_target - k)                                  # voxels of the target level one node of level k covers
            lo = [pos[a] * span for a in range(3)]
            if any(lo[a] >= dims[a] for a in range(3)) or not node.in_volume:
                return
            hi = [min(lo[a] + span, dims[a]) for a in range(3)]
            if k == k_target or not node.children:
                if k == k_target and node.brick is not None:
                    vol[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]] = node.brick[:hi[2] - lo[2], :hi[1] - lo[1], :hi[0] - lo[0]]
                else:
                    vol[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]] = node.avg
                return
            for i, c in enumerate(node.children):
                visit(c, k + 1, (pos[0] * 2 + (i & 1), pos[1] * 2 + ((i >> 1) & 1), pos[2] * 2 + (i >> 2)))
        visit(self.root, 0, (0, 0, 0))
        return vol


def mip_map(vol, levels):
    """Reference mip map of the test (octreetest.cpp:322-334, VolumeOperatorHalfsample on uint16): each level halves the previous one
    (ceil for odd sizes) by averaging the 2x2x2 block, the last voxel replicated where the block leaves the volume."""
    out = [to_uint16(vol)]
    for _ in range(1, levels):
        p = out[-1].astype(np.uint64)
        dz, dy, dx = p.shape
        nz, ny, nx = -(-dz // 2), -(-dy // 2), -(-dx // 2)
        iz = np.minimum(np.arange(nz * 2), dz - 1); iy = np.minimum(np.arange(ny * 2), dy - 1); ix = np.minimum(np.arange(nx * 2), dx - 1)
        q = p[np.ix_(iz, iy, ix)]
        s = q.reshape(nz, 2, ny, 2, nx, 2).sum(axis=(1, 3, 5))
        out.append((s // 8).astype(np.uint16))
    return out


# the reference's deterministic test volumes: RandomNumberGenerator (octreetest.cpp:174-192), createNoiseVolume (:201-208), createBrickedVolume (:219-262)
class RandomNumberGenerator:
    def __init__(self, mz=3, mw=17):
        self.mz, self.mw = mz, mw

    def get_uint32(self):
        self.mz = (36969 * (self.mz & 65535) + (self.mz >> 16)) & 0xffffffff
        self.mw = (18000 * (self.mw & 65535) + (self.mw >> 16)) & 0xffffffff
        return ((self.mz << 16) + self.mw) & 0xffffffff

    def get_uint16(self):
        return self.get_uint32() >> 16


def noise_volume(dims, mz=3, mw=17):
    rng = RandomNumberGenerator(mz, mw)
    n = dims[0] * dims[1] * dims[2]
    return np.array([rng.get_uint16() for _ in range(n)], np.uint16).reshape(dims[::-1])


def bricked_volume(dims, brick, mz=3, mw=17, min_diff=1 << 13):
    """createBrickedVolume (octreetest.cpp:219-262): homogeneous regions whose values differ from their -x / -y / -z neighbours."""
    rng = RandomNumberGenerator(mz, mw)
    nb = [dims[a] // brick[a] for a in range(3)]
    vals = np.zeros(nb[::-1], np.int64)
    vol = np.zeros(dims[::-1], np.uint16)
    for z in range(nb[2]):
        for y in range(nb[1]):
            for x in range(nb[0]):
                while True:
                    v = rng.get_uint16()
                    if x > 0 and abs(v - vals[z, y, x - 1]) < min_diff:
                        continue
                    if y > 0 and abs(v - vals[z, y - 1, x]) < min_diff:
                        continue
                    if z > 0 and abs(v - vals[z - 1, y, x]) < min_diff:
                        continue
                    break
                vals[z, y, x] = v
                vol[z * brick[2]:(z + 1) * brick[2], y * brick[1]:(y + 1) * brick[1], x * brick[0]:(x + 1) * brick[0]] = v
    return vol


def pool_arrays(tree):
    """The product's layout: all levels concatenated (offset (8^k - 1) / 7), node index = offset + (z * G + y) * G + x;
    returns avg, mn, mx (uint16), flags (uint8) and the brick pool (nodes, B, B, B) uint16 (zeros where a node has no brick)."""
    d = tree.dense()
    B = tree.B
    total = sum(8 ** k for k in range(tree.num_levels))
    avg = np.zeros(total, np.uint16); mn = np.zeros(total, np.uint16); mx = np.zeros(total, np.uint16); flags = np.zeros(total, np.uint8)
    pool = np.zeros((total, B, B, B), np.uint16)
    for k, lv in enumerate(d):
        off, G = (8 ** k - 1) // 7, 1 << k
        avg[off:off + G ** 3] = lv["avg"].reshape(-1); mn[off:off + G ** 3] = lv["mn"].reshape(-1); mx[off:off + G ** 3] = lv["mx"].reshape(-1)
        flags[off:off + G ** 3] = lv["flags"].reshape(-1)
        for (x, y, z), b in lv["bricks"].items():
            pool[off + (z * G + y) * G + x] = b
    return avg, mn, mx, flags, pool


# ---- OptimizedProxyGeometry "volumeoctree" mode -----------------------------------------------------------------------------------------
def proxy_boxes(tree, region_empty, clip=None):
    """VolumeOctreeBackgroundThread::traverseOctreeAndCreateGeometry (V/modules/base/processors/proxygeometry/
    optimizedproxygeometry.cpp:1657-1737) over the pointer tree: one cube (llf, urb in TEXTURE space, as addCubeMesh :480-491 emits
    them) per leaf whose [min, max] (uint16) `region_empty(mn, mx)` does not call transparent.  clip = (llf, urb) voxel clip bounds or None."""
    dims = np.array(tree.dims, np.float32)
    out = []
    f32 = np.float32

    def visit(node, llf, urb):
        if not node.children:
            if region_empty(node.mn, node.mx):
                return
            if np.any(llf >= dims):                                                   # :1684-1685
                return
            u = np.minimum(urb, dims - f32(1))
            lo = np.maximum(llf - f32(1), f32(-0.5))                                  # :1693-1694
            hi = np.minimum(u + f32(1), dims - f32(0.5))
            if clip is not None:
                c_lo, c_hi = np.array(clip[0], np.float32) - f32(0.5), np.array(clip[1], np.float32) + f32(0.5)
                if not (np.all(lo >= c_lo) and np.all(hi <= c_hi)):                   # not containsVolume
                    if np.any(lo > c_hi) or np.any(hi < c_lo):                        # not intersects either (touching counts, bounds.h:162-170)
                        return
                    lo, hi = np.maximum(lo, c_lo), np.minimum(hi, c_hi)               # addCubeMeshClip (:493-512)
            out.append(np.concatenate([(lo + f32(0.5)) / dims, (hi + f32(0.5)) / dims]).astype(np.float32))
            return
        half = (urb - llf + f32(1)) / f32(2)
        for i, c in enumerate(node.children):
            if c is None:
                continue
            cl = llf + np.array([i & 1, (i >> 1) & 1, i >> 2], np.float32) * half
            visit(c, cl, cl + half - f32(1))
    visit(tree.root, np.zeros(3, np.float32), np.full(3, tree.octree_dim - 1, np.float32))
    return np.array(out, np.float32).reshape(-1, 6)
