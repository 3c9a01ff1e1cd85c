"""Fluent `.cas` import: mirror of the reference's `importers.FluentCase` (I/FluentReader.cpp) for
ASCII and binary case files with one cell zone -- the meshes of the reference's own thermal / flow tests
(T/CellMark/cav32.cas, T/3d-cube.cas, ...). What matters downstream is the NUMBERING the reference
reader produces, because it fixes the face order of every cell and with it the floating-point
summation order of the assembly:

* faces keep their file index; boundary faces get a ghost cell numbered numCells + (running count in
  file order) (I/FluentReader.cpp:323-348); the interior cell always comes first in faceCells, and
  the node order of a face is reversed iff (dimension == 3) xor (c0 == 0) (:313-356);
* a face zone is classified by its first face (I/FluentReader.cpp:656-692); the mesh lists the faces
  of the interior zones first, then the boundary zones, each in ascending zone id, and makes one
  face group per boundary zone with id = zone id and groupType = the zone's type string
  ("wall", "velocity-inlet", ...; :696-760);
* cells: the zone's cells in file order, then the ghost cells in the order of those boundary faces
  (:795-823).
Node numbering is kept global (the reference renumbers nodes per mesh; geometry does not depend on it).
"""
import re

import numpy as np

from . import meshgen
from .models import CException, Mesh

_HEX = r"([0-9a-fA-F]+)"


class FluentCase:
    """`fc = FluentCase(path); fc.read(); meshes = fc.getMeshList()` (I/FluentReader.h:94-150)."""

    def __init__(self, fileName):
        self.fileName = fileName
        self._read = False

    # ------------------------------------------------------------------ parsing
    @staticmethod
    def _close(buf, i):
        """index just past the parenthesis that closes the one at buf[i] (strings may hold parentheses)"""
        depth, n, in_str = 0, len(buf), False
        OPEN, CLOSE, QUOTE = 40, 41, 34
        while i < n:
            c = buf[i]
            if in_str:
                if c == QUOTE:
                    in_str = False
            elif c == QUOTE:
                in_str = True
            elif c == OPEN:
                depth += 1
            elif c == CLOSE:
                depth -= 1
                if depth == 0:
                    return i + 1
            i += 1
        raise CException("FluentCase: unbalanced parentheses")

    @staticmethod
    def _skip_ws(buf, k):
        n = len(buf)
        while k < n and buf[k] in b" \t\r\n":
            k += 1
        return k

    def _end_binary(self, buf, k, sid):
        """closeSectionBinary (I/FluentReader.cpp): the payload is followed by `)` and the line
        `End of Binary Section   <id>)`"""
        m = re.compile(rb"End of Binary Section\s+%d\s*\)" % sid).search(buf, k, k + 256)
        if not m:
            raise CException("FluentCase: binary section %d is not terminated where its header says" % sid)
        return m.end()

    def read(self):
        try:
            self._read_sections()
        except (ValueError, IndexError) as e:   # a payload shorter than its header says, a non-numeric token ...
            raise CException("FluentCase: %s is damaged or truncated (%s)" % (self.fileName, e))

    def _read_sections(self):
        with open(self.fileName, "rb") as fh:
            buf = fh.read()
        self._dim = None
        self._num_nodes = self._num_cells = self._num_faces = 0
        self._coords = None
        self._cell_zones = {}        # id -> (iBeg, iEnd, type)
        self._face_zones = {}        # id -> dict(iBeg, iEnd, type)
        self._zone_types = {}        # id -> type string of sections 39 / 45
        face_blocks = []             # (zoneId, iBeg, iEnd, shape, ints) in file order
        head = re.compile(rb"\(\s*(\d+)\s*")
        hdr = re.compile(rb"\(\s*\d+\s*\(\s*([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s+([0-9a-fA-F]+)\s*"
                         rb"([0-9a-fA-F]*)\s*\)")
        zone_re = re.compile(rb"\(\s*\d+\s*\(\s*(\d+)\s+([^\s()]+)\s+([^\s()]+)")
        i, n = 0, len(buf)
        while True:
            i = buf.find(b"(", i)
            if i < 0:
                break
            m = head.match(buf, i, i + 32)
            if not m:
                i = self._close(buf, i)
                continue
            sid = int(m.group(1))
            binary, dp = sid > 1000, sid > 3000      # I/FluentReader.cpp:428-429
            kind = sid % 1000
            if kind == 2:
                end = self._close(buf, i)
                self._dim = int(buf[m.end():end - 1])
                i = end
            elif kind in (10, 12, 13):
                h = hdr.match(buf, i, i + 160)
                if not h:
                    raise CException("FluentCase: cannot read the header of a section %d" % sid)
                zone, beg, end_, typ = (int(h.group(k), 16) for k in (1, 2, 3, 4))
                last = int(h.group(5), 16) if h.group(5) else 0
                count = end_ - beg + 1
                k = self._skip_ws(buf, h.end())
                values = None
                if k < n and buf[k] == 40:   # a data list follows
                    if binary:
                        p0 = k + 1
                        if kind == 10:
                            dim = last or self._dim
                            nbytes = count * dim * (8 if dp else 4)
                            values = np.frombuffer(buf, dtype="<f8" if dp else "<f4", count=count * dim, offset=p0)
                            values = values.astype(np.float64).reshape(-1, dim)
                        elif kind == 13:
                            shape = last if last >= 0 else self._dim
                            if shape in (0, 5):   # mixed: walk the records
                                q, recs = p0, []
                                for _ in range(count):
                                    nn = int(np.frombuffer(buf, dtype="<i4", count=1, offset=q)[0])
                                    recs.append(np.frombuffer(buf, dtype="<i4", count=nn + 3, offset=q))
                                    q += 4 * (nn + 3)
                                values = np.concatenate(recs).astype(np.int64)
                                nbytes = q - p0
                            else:
                                nbytes = count * (shape + 2) * 4
                                values = np.frombuffer(buf, dtype="<i4", count=count * (shape + 2), offset=p0).astype(np.int64)
                        else:
                            nbytes = count * 4    # mixed cell zone: one element type per cell
                        i = self._end_binary(buf, p0 + nbytes, sid)
                    else:
                        dend = self._close(buf, k)
                        data = buf[k + 1:dend - 1]
                        if kind == 10:
                            dim = last or self._dim
                            values = np.array(data.split(), dtype=np.float64).reshape(-1, dim)
                        elif kind == 13:
                            values = np.array([int(t, 16) for t in data.split()], dtype=np.int64)
                        k = self._skip_ws(buf, dend)
                        if k >= n or buf[k] != 41:
                            raise CException("FluentCase: malformed section %d" % sid)
                        i = k + 1
                elif binary:
                    i = self._end_binary(buf, k, sid)
                else:
                    if k >= n or buf[k] != 41:
                        raise CException("FluentCase: malformed section %d" % sid)
                    i = k + 1
                if kind == 10:
                    if zone == 0:
                        self._num_nodes = end_
                    elif values is not None:
                        if self._coords is None:
                            self._coords = np.zeros((self._num_nodes, 3))
                        self._coords[beg - 1:end_, :values.shape[1]] = values
                elif kind == 12:
                    if zone == 0:
                        self._num_cells = end_
                    elif typ in (1, 17):
                        self._cell_zones[zone] = (beg - 1, end_ - 1, typ)
                    elif typ == 32:
                        self._num_cells -= count
                    else:
                        raise CException("cell thread type not handled")
                else:
                    if zone == 0:
                        self._num_faces = end_
                    elif typ not in (0, 31):
                        if values is None:
                            raise CException("FluentCase: face zone %d has no data" % zone)
                        self._face_zones[zone] = dict(iBeg=beg - 1, iEnd=end_ - 1, type=typ)
                        face_blocks.append((zone, beg - 1, end_ - 1, last, values))
                    else:
                        self._num_faces -= count
            elif kind in (39, 45) and not binary:
                h = zone_re.match(buf, i, i + 256)
                if h:
                    self._zone_types[int(h.group(1))] = h.group(2).decode("latin-1")
                i = self._close(buf, i)
            elif binary:   # a binary section this reader does not use (node flags, cell trees, ...): skip to its end mark
                mm = re.compile(rb"End of Binary Section\s+%d\s*\)" % sid).search(buf, i)
                if not mm:
                    raise CException("FluentCase: binary section %d has no end mark" % sid)
                i = mm.end()
            else:
                i = self._close(buf, i)
        if self._dim is None or self._coords is None or not face_blocks:
            raise CException("FluentCase: %s holds no mesh" % self.fileName)
        # ---- faces in file order: nodes, cells, ghost numbering
        nf = self._num_faces
        dim = self._dim
        self._fc = np.full((nf, 2), -1, np.int64)
        self._fn = [None] * nf
        ghosts = 0
        for zone, beg, end_, shape, ints in face_blocks:
            if shape < 0:
                shape = dim
            p = 0
            for f in range(beg, end_ + 1):
                nn = shape
                if shape in (0, 5):
                    nn = int(ints[p]); p += 1
                nodes = ints[p:p + nn] - 1
                c0, c1 = int(ints[p + nn]), int(ints[p + nn + 1])
                p += nn + 2
                if c0 == 0 and c1 == 0:
                    raise CException("FluentCase: boundary meshes without cells are not supported")
                reverse = dim == 3
                if c0 == 0:
                    reverse = not reverse
                cells = [c - 1 for c in (c0, c1) if c != 0]
                if len(cells) == 1:
                    cells.append(self._num_cells + ghosts)
                    ghosts += 1
                self._fc[f] = cells
                self._fn[f] = nodes[::-1].copy() if reverse else nodes.copy()
        self._num_boundary_faces = ghosts
        self._read = True

    # ------------------------------------------------------------------ meshes
    def getMeshList(self):
        if not self._read:
            raise CException("FluentCase: call read() first")
        if len(self._cell_zones) != 1:
            raise CException("FluentCase: %d cell zones -- only single-zone case files are supported" % len(self._cell_zones))
        (czid, (cbeg, cend, _)), = self._cell_zones.items()
        ncell = self._num_cells
        interior, boundary = [], []
        for zid in sorted(self._face_zones):           # std::map order
            z = self._face_zones[zid]
            c1 = self._fc[z["iBeg"], 1]
            (boundary if c1 >= ncell else interior).append(zid)
        face_list, sizes, ids, types = [], [], [0], ["interior"]
        for zid in interior:
            z = self._face_zones[zid]
            face_list.extend(range(z["iBeg"], z["iEnd"] + 1))
        sizes.append(len(face_list))
        boundary_cells = []
        for zid in boundary:
            z = self._face_zones[zid]
            rng = range(z["iBeg"], z["iEnd"] + 1)
            face_list.extend(rng)
            sizes.append(len(rng))
            ids.append(zid)
            types.append(self._zone_types.get(zid, "wall"))
            boundary_cells.extend(int(self._fc[f, 1]) for f in rng)
        n_mesh_cells = cend - cbeg + 1
        g2l = {}
        for k, c in enumerate(range(cbeg, cend + 1)):
            g2l[c] = k
        for k, c in enumerate(boundary_cells):
            g2l[c] = n_mesh_cells + k
        fc = np.array([[g2l[int(self._fc[f, 0])], g2l[int(self._fc[f, 1])]] for f in face_list], np.int32)
        counts = np.array([len(self._fn[f]) for f in face_list], np.int32)
        fnodes = np.concatenate([self._fn[f] for f in face_list]).astype(np.int32)
        raw = meshgen._finish(self._dim, n_mesh_cells, self._coords, fc, fnodes, counts, sizes)
        raw.group_id = np.array(ids, np.int32)
        raw.group_types = types
        raw.group_kind = np.array([0] + [3 if t == "symmetry" else 1 for t in types[1:]], np.int32)
        raw.cell_zone_id = czid
        mesh = Mesh(raw, group_types=types)
        return [mesh]
