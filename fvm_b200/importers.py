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
        self._zone_names, self._zone_vars, self._rp_vars = {}, {}, ""
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
                end = self._close(buf, i)
                h = zone_re.match(buf, i, i + 256)
                if h:
                    zid = int(h.group(1))
                    self._zone_types[zid] = h.group(2).decode("latin-1")
                    self._zone_names[zid] = h.group(3).decode("latin-1")
                    k = self._skip_ws(buf, self._close(buf, buf.find(b"(", i + 1)))   # past the header list
                    if k < end and buf[k] == 40:
                        self._zone_vars[zid] = buf[k:self._close(buf, k)].decode("latin-1")
                i = end
            elif kind == 37 and not binary:
                end = self._close(buf, i)
                k = buf.find(b"(", i + 1)
                self._rp_vars = buf[k:self._close(buf, k)].decode("latin-1")
                i = end
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
        # Nodes of the mesh in the reference's order (I/FluentReader.cpp:841-856: cellNodes = cellFaces x faceNodes of the
        # whole file, localized over the zone's cells): numbered as the zone's cells meet them, cell after cell, a cell's
        # nodes in the order of its faces (file order) and of their stored node lists. Nodes no cell of the zone uses
        # are dropped. (No kernel depends on it; the Tecplot dumps of the reference's test scripts do.)
        faces_of = [[] for _ in range(cend + 1)]
        for f in range(self._num_faces):
            for c in self._fc[f]:
                if cbeg <= c <= cend:
                    faces_of[int(c)].append(f)
        new_id = np.full(len(self._coords), -1, np.int64)
        order = []
        for c in range(cbeg, cend + 1):
            for f in faces_of[c]:
                for v in self._fn[f]:
                    if new_id[v] < 0:
                        new_id[v] = len(order)
                        order.append(int(v))
        fnodes = new_id[fnodes].astype(np.int32)
        raw = meshgen._finish(self._dim, n_mesh_cells, self._coords[np.asarray(order, np.int64)], fc, fnodes, counts, sizes)
        raw.group_id = np.array(ids, np.int32)
        raw.group_types = types
        raw.group_kind = np.array([0] + [3 if t == "symmetry" else 1 for t in types[1:]], np.int32)
        raw.cell_zone_id = czid
        mesh = Mesh(raw, group_types=types)
        self._mesh_of_zone = {czid: mesh}
        self._interior_zone_ids = {czid: interior}
        return [mesh]

    def getCellZones(self):
        """{id: zone} with .ID, .iBeg, .iEnd (0-based, inclusive), .mesh, .interiorZoneIds (I/FluentReader.h:30-60)"""
        out = {}
        for zid, (beg, end, _) in self._cell_zones.items():
            z = type("FluentCellZone", (), {})()
            z.ID, z.iBeg, z.iEnd = zid, beg, end
            z.mesh = getattr(self, "_mesh_of_zone", {}).get(zid)
            z.interiorZoneIds = list(getattr(self, "_interior_zone_ids", {}).get(zid, []))
            out[zid] = z
        return out

    def getFaceZones(self):
        out = {}
        for zid, d in self._face_zones.items():
            z = type("FluentFaceZone", (), {})()
            z.ID, z.iBeg, z.iEnd, z.threadType = zid, d["iBeg"], d["iEnd"], d["type"]
            z.zoneType, z.zoneName = self._zone_types.get(zid, ""), self._zone_names.get(zid, "")
            out[zid] = z
        return out

    # ------------------------------------------------------------------ case variables / boundary conditions
    def getVar(self, name):
        """rp-variable of the case file (section 37), e.g. 'mom/relax' (scripts/FluentCase.py:165-183)"""
        return self._vars()[name]

    def _vars(self):
        if not hasattr(self, "_vars_dict"):
            self._vars_dict = _alist(_parse_scheme(self._rp_vars)) if self._rp_vars else {}
        return self._vars_dict

    def _zone(self, zid):
        if not hasattr(self, "_zones"):
            self._zones = {}
        if zid not in self._zones:
            txt = self._zone_vars.get(zid, "")
            self._zones[zid] = _FluentZone(zid, self._zone_names.get(zid, ""), self._zone_types.get(zid, ""),
                                           _alist(_parse_scheme(txt)) if txt else {})
        return self._zones[zid]

    def importThermalBCs(self, tmodel):
        """scripts/FluentCase.py:218-249"""
        for gid, bc in tmodel.getBCMap().items():
            z = self._zone(gid)
            if z.zoneType == "wall":
                kind = z.getVar("thermal-bc")
                if kind == 0:
                    bc.bcType = "SpecifiedTemperature"
                    bc.setVar("specifiedTemperature", z.getConstantVar("t"))
                elif kind == 1:
                    bc.bcType = "SpecifiedHeatFlux"
                    bc.setVar("specifiedHeatFlux", z.getConstantVar("q"))
                elif kind == 3:
                    bc.bcType = "CoupledWall"
                else:
                    raise TypeError("thermal BCType %d not handled" % kind)
            elif z.zoneType in ("velocity-inlet", "pressure-inlet", "pressure-outlet", "mass-flow-inlet", "exhaust-fan",
                                "intake-fan", "inlet-vent", "outlet-vent"):
                bc.bcType = "SpecifiedTemperature"
                bc.setVar("specifiedTemperature", z.getConstantVar("t" if z.zoneType == "velocity-inlet" else "t0"))
            elif z.zoneType == "symmetry":
                pass
            else:
                raise TypeError("invalid boundary type : " + z.zoneType)

    def importFlowBCs(self, fmodel, meshes):
        """scripts/FluentCase.py:251-318: initial values and under-relaxation factors from the case variables,
        one boundary condition per face zone, density / viscosity from the cell zone's material."""
        o = fmodel.getOptions()
        o["initialXVelocity"] = self.getVar("x-velocity/default")
        o["initialYVelocity"] = self.getVar("y-velocity/default")
        o["initialZVelocity"] = self.getVar("z-velocity/default")
        o["initialPressure"] = self.getVar("pressure/default")
        o["momentumURF"] = self.getVar("mom/relax")
        o["pressureURF"] = self.getVar("pressure/relax")
        for gid, bc in fmodel.getBCMap().items():
            z = self._zone(gid)
            if z.zoneType == "wall":
                motion = z.getVar("motion-bc")
                bc.bcType = "NoSlipWall"
                if motion == 1:
                    vmag = z.getVar("vmag")
                    bc["specifiedXVelocity"] = vmag * z.getConstantVar("ni")
                    bc["specifiedYVelocity"] = vmag * z.getConstantVar("nj")
                    bc["specifiedZVelocity"] = vmag * z.getConstantVar("nk")
                elif motion != 0:
                    raise TypeError("flow BCType %d not handled" % motion)
            elif z.zoneType == "velocity-inlet":
                spec = z.getVar("velocity-spec")
                bc.bcType = "VelocityBoundary"
                if spec == 0:
                    vmag = z.getVar("vmag")
                    bc["specifiedXVelocity"] = vmag * z.getConstantVar("ni")
                    bc["specifiedYVelocity"] = vmag * z.getConstantVar("nj")
                    bc["specifiedZVelocity"] = vmag * z.getConstantVar("nk")
                elif spec == 1:
                    bc["specifiedXVelocity"] = z.getConstantVar("u")
                    bc["specifiedYVelocity"] = z.getConstantVar("v")
                    bc["specifiedZVelocity"] = z.getConstantVar("w")
                else:
                    raise TypeError("flow BCType %d not handled" % spec)
            elif z.zoneType == "pressure-outlet":
                bc.bcType = "PressureBoundary"
                bc["specifiedPressure"] = z.getConstantVar("p")
            elif z.zoneType == "pressure-inlet":
                bc.bcType = "PressureBoundary"
                bc["specifiedPressure"] = z.getConstantVar("p0")
            elif z.zoneType == "symmetry":
                bc.bcType = "Symmetry"
            else:
                raise TypeError("invalid boundary type : " + z.zoneType)
        materials = _alist(self.getVar("materials")) if isinstance(self.getVar("materials"), list) else {}
        for mesh in meshes:
            vc = fmodel.getVCMap()[mesh.getID()]
            cz = self._zone(mesh.raw.cell_zone_id)
            mat = materials[cz.getVar("material")]
            props = _alist(mat[1:]) if isinstance(mat, list) else {}
            vc["density"] = _constant(props["density"])
            vc["viscosity"] = _constant(props["viscosity"])


class _FluentZone:
    """scripts/FluentCase.py:86-130"""

    def __init__(self, zid, name, zone_type, vars_):
        self.id, self.zoneName, self.zoneType, self.varsDict = zid, name or "%s_%d" % (zone_type, zid), zone_type, vars_

    def getVar(self, v):
        return self.varsDict[v]

    def getConstantVar(self, v):
        return _constant(self.varsDict[v])


def _constant(val):
    """(name (constant . 300) (profile "" "")) -> 300"""
    if not isinstance(val, list):
        return val
    if val and isinstance(val[0], list):
        val = val[0]
    if val and val[0] == "constant":
        return val[1]
    raise ValueError("value is not constant: %r" % (val,))


def _parse_scheme(text):
    """Scheme data -> nested Python lists; a dotted pair (a . b) becomes [a, b]; #t / #f -> bool; numbers -> int / float;
    strings and symbols -> str. Enough for the association lists Fluent writes into a case file."""
    tok = re.findall(r'"(?:[^"\\]|\\.)*"|[()]|[^\s()"]+', text)
    pos = 0

    def atom(t):
        if t[0] == '"':
            return t[1:-1]
        if t == "#t":
            return True
        if t == "#f":
            return False
        try:
            return int(t)
        except ValueError:
            try:
                return float(t)
            except ValueError:
                return t

    def parse():
        nonlocal pos
        t = tok[pos]
        pos += 1
        if t != "(":
            return atom(t)
        out = []
        while tok[pos] != ")":
            if tok[pos] == ".":
                pos += 1
                continue
            out.append(parse())
        pos += 1
        return out

    return parse() if tok else []


def _alist(items):
    """[[key, value...], ...] -> {key: value}: one value stays itself, several stay a list"""
    d = {}
    for it in items:
        if isinstance(it, list) and it and isinstance(it[0], str):
            rest = it[1:]
            d[it[0]] = rest[0] if len(rest) == 1 else rest
    return d


class MMReader:
    """MatrixMarket matrix + right-hand-side pair -> one scalar linear system (I/MMReader.cpp:24-184; the input of the
    reference's testLinearSolver, T/TESTS Fvm001). Off-diagonal entries keep their file order inside each row (that is
    the order every row sum runs in), the diagonal is stored apart, symmetric files are expanded, and b = -rhs
    (the library's sign convention r = b + A x). `getLS(lib)` returns a `capi.DeviceSystem` ready for any solver."""

    def __init__(self, matrixFileName, rhsFileName):
        self.matrixFileName, self.rhsFileName = matrixFileName, rhsFileName

    def read(self):
        with open(self.matrixFileName) as fh:
            tokens = fh.read().split()
        if tokens[0] == "%%" and tokens[1] == "MatrixMarket":
            tokens = tokens[2:]
        elif tokens[0] == "%%MatrixMarket":
            tokens = tokens[1:]
        else:
            raise CException("not a MatrixMarket file")
        mtype, coord, ftype, symm = tokens[:4]
        if mtype != "matrix":
            raise CException("not a MatrixMarket file")
        if coord != "coordinate":
            raise CException("not a sparse matrix")
        if ftype != "real":
            raise CException("not a real matrix")
        if symm not in ("symmetric", "general"):
            raise CException("not symmetric or general matrix")
        n, ncol, nnz = (int(t) for t in tokens[4:7])
        if n != ncol:
            raise CException("not a square matrix")
        body = tokens[7:7 + 3 * nnz]
        rows = [[] for _ in range(n)]
        diag = np.zeros(n)
        for e in range(nnz):
            i, j, c = int(body[3 * e]) - 1, int(body[3 * e + 1]) - 1, float(body[3 * e + 2])
            if i != j:
                rows[i].append((j, c))
                if symm == "symmetric":
                    rows[j].append((i, c))
            else:
                diag[i] = c
        row = np.zeros(n + 1, np.int32)
        row[1:] = np.cumsum([len(r) for r in rows])
        col = np.array([j for r in rows for j, _ in r], np.int32)
        off = np.array([c for r in rows for _, c in r], np.float64)
        with open(self.rhsFileName) as fh:
            b = -np.array(fh.read().split()[:n], dtype=np.float64)
        return dict(n=n, row=row, col=col, diag=diag, off=off, b=b)

    def getLS(self, lib):
        from . import capi
        d = self.read()
        return capi.DeviceSystem(lib, raw=(d["n"], 0, d["row"], d["col"], d["diag"], d["off"], d["b"]))
