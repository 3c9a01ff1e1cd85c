"""Fluent `.dat` export: mirror of the reference's `exporters.FluentDataExporterA`
(src/fvm/src/modules/exporters/FluentDataExporter.h:14-187, text mode) -- the writer behind the reference's
registered FlowModel test (T/TESTS Fvm002: FvmTestFlowModel.py cav32 -> cav32-prism.dat). Cell fields are
written per cell zone and then per boundary face group (the ghost cells' values), face fields per interior
face zone and per boundary face group, every value as "%12.5e" (F/ArrayWriter.h:32)."""
import numpy as np

from .models import CException


class FluentDataExporterA:
    def __init__(self, reader, fileName, binary=False, atypeComponent=0):
        if binary:
            raise CException("FluentDataExporter: binary output is not supported")
        self._reader, self._section = reader, 300
        try:
            self._fp = open(fileName, "w", newline="")
        except OSError:
            raise CException("FluentDataExporter: cannot open file " + fileName + "for writing")

    def init(self):
        self._fp.write("(4 (60 0 0 1 2 4 4 4 8 8 4))\n")

    def _block(self, field_id, zone_id, beg, end, values):
        self._fp.write("(%d (%d %d 1 0 1 %d %d)\n(" % (self._section, field_id, zone_id, beg + 1, end + 1))
        self._fp.write("".join("%12.5e\n" % v for v in values))
        self._fp.write("))\n")

    def _write(self, field, field_id, component=None):
        face_zones, cell_zones = self._reader.getFaceZones(), self._reader.getCellZones()
        for czid in sorted(cell_zones):
            cz = cell_zones[czid]
            mesh = cz.mesh
            if mesh is None:
                raise CException("FluentDataExporter: call getMeshList() on the reader first")
            cells, faces = mesh.getCells(), mesh.getFaces()
            if cells in field:
                a = np.asarray(field[cells])
                if component is not None:
                    a = a[:, component]
                self._block(field_id, cz.ID, cz.iBeg, cz.iEnd, a[:cells.getSelfCount()])
                for fg in mesh.getBoundaryFaceGroups():
                    fz = face_zones[fg.id]
                    off, cnt = fg.site.getOffset(), fg.site.getCount()
                    cbeg = int(mesh.raw.face_cells[off, 1])      # the group's ghost cells are contiguous
                    self._block(field_id, fz.ID, fz.iBeg, fz.iEnd, a[cbeg:cbeg + cnt])
            if component is None and faces in field:
                a = np.asarray(field[faces])
                off = 0
                for fzid in cz.interiorZoneIds:
                    fz = face_zones[fzid]
                    cnt = fz.iEnd - fz.iBeg + 1
                    self._block(field_id, fz.ID, fz.iBeg, fz.iEnd, a[off:off + cnt])
                    off += cnt
                for fg in mesh.getBoundaryFaceGroups():
                    fz = face_zones[fg.id]
                    o, cnt = fg.site.getOffset(), fg.site.getCount()
                    self._block(field_id, fz.ID, fz.iBeg, fz.iEnd, a[o:o + cnt])

    def writeScalarField(self, field, fluentFieldId):
        self._write(field, fluentFieldId)

    def writeVectorField(self, field, fluentFieldId):
        for nd in range(3):
            self._write(field, fluentFieldId + nd, component=nd)

    def finish(self):
        self._fp.close()


# ---------------------------------------------------------------------------------------------------------------
# Tecplot point of view of a mesh: the cell -> node connectivity in the element's canonical node order, and the
# finite-element zone file the reference's test scripts write with it (dumpTecplotFile in
# T/THERMAL_MATRIX/testThermalParallel.py:54-150 and its siblings: `temp_procN.dat`).

# canonical faces of the element types as lists of canonical node numbers (the conventions of F/Cell.cpp:20-50; face 0
# lists nodes 0, 1, 2[, 3] in this order for every type, which the ordering below relies on)
_ELEMENT_FACES = {
    "quad": [(0, 1), (1, 2), (2, 3), (3, 0)],
    "tri": [(0, 1), (1, 2), (2, 0)],
    "hexa": [(0, 1, 2, 3), (4, 7, 6, 5), (0, 4, 5, 1), (1, 5, 6, 2), (2, 6, 7, 3), (3, 7, 4, 0)],
    "tetra": [(0, 1, 2), (0, 3, 1), (1, 3, 2), (2, 3, 0)],
    "pyramid": [(0, 1, 2, 3), (0, 4, 1), (1, 4, 2), (2, 4, 3), (3, 4, 0)],
    "prism": [(0, 1, 2), (3, 5, 4), (0, 3, 4, 1), (1, 4, 5, 2), (2, 5, 3, 0)],
}
TECPLOT_ZONE_TYPE = {"tri": "FETRIANGLE", "quad": "FEQUADRILATERAL", "tetra": "FETETRAHEDRON", "hexa": "FEBRICK"}


def _element_type(n_nodes, face_sizes):
    e, t, q = face_sizes.count(2), face_sizes.count(3), face_sizes.count(4)
    if n_nodes == 4 and e == 4:
        return "quad"
    if n_nodes == 3 and e == 3:
        return "tri"
    if n_nodes == 8 and q == 6:
        return "hexa"
    if n_nodes == 4 and t == 4:
        return "tetra"
    if n_nodes == 5 and t == 4 and q == 1:
        return "pyramid"
    if n_nodes == 6 and t == 2 and q == 3:
        return "prism"
    return None


def cell_nodes(raw):
    """Mesh::getCellNodes() (F/Mesh.cpp:425-451) for the self cells of a raw mesh: (row, col) with the nodes of every
    cell in the canonical order of its element type (Cell<T>::orderCellFacesAndNodes, F/Cell.cpp:96-200).

    The cell's first face (lowest face index) with as many nodes as the element's canonical face 0 is laid onto that
    canonical face -- its node list as stored if the cell is the face's c0, reversed if it is c1 --; every other face
    is identified with the canonical face that shares the same subset of those nodes; a node then is the canonical
    node common to exactly the canonical faces it belongs to."""
    fc = np.asarray(raw.face_cells).reshape(-1, 2)
    fnc = np.asarray(raw.face_node_count)
    fstart = np.concatenate([[0], np.cumsum(fnc)])
    fn = np.asarray(raw.face_nodes)
    n = int(raw.n_cells)
    faces_of = [[] for _ in range(n)]
    for f in range(len(fc)):                       # ascending face index per cell, both sides (the transpose of faceCells)
        for c in (int(fc[f, 0]), int(fc[f, 1])):
            if c < n:
                faces_of[c].append(f)
    row, col = np.zeros(n + 1, np.int32), []
    for c in range(n):
        faces = faces_of[c]
        nodes_of = [fn[fstart[f]:fstart[f + 1]].tolist() for f in faces]
        all_nodes = sorted(set(v for nl in nodes_of for v in nl))
        kind = _element_type(len(all_nodes), [len(nl) for nl in nodes_of])
        if kind is None:
            raise CException("cell_nodes: unsupported element (cell %d: %d nodes)" % (c, len(all_nodes)))
        tmpl = _ELEMENT_FACES[kind]
        k0 = next(k for k, nl in enumerate(nodes_of) if len(nl) == len(tmpl[0]))
        first = nodes_of[k0] if int(fc[faces[k0], 0]) == c else nodes_of[k0][::-1]
        bit = {v: 1 << i for i, v in enumerate(first)}
        canon = {sum(1 << v for v in t if v in tmpl[0]): i for i, t in enumerate(tmpl)}   # subset of face 0 -> canonical face
        mask = {}
        for nl in nodes_of:
            i = canon[sum(bit.get(v, 0) for v in nl)]
            m = sum(1 << v for v in tmpl[i])
            for v in nl:
                mask[v] = mask.get(v, ~0) & m
        ordered = [0] * len(all_nodes)
        for v, m in mask.items():
            ordered[m.bit_length() - 1] = v
        col.extend(ordered)
        row[c + 1] = len(col)
    return row, np.asarray(col, np.int32)


def _py2_str(x):
    """str() of a float as the reference's Python 2 scripts print it: 12 significant digits, always a float."""
    s = "%.12g" % float(x)
    return s if any(ch in s for ch in ".en") else s + ".0"


def dumpTecplotFile(fileName, meshes, mtype, cellField, geomFields, title=" tecplot file for 2D Cavity problem ",
                    varName="velX"):
    """The finite-element Tecplot file of the reference's parallel test scripts (dumpTecplotFile,
    T/THERMAL_MATRIX/testThermalParallel.py:54-150): one zone per mesh, node coordinates in BLOCK packing, the cell
    field and the cells' centroid y as cell-centred variables, then the 1-based cell -> node connectivity.
    mtype: 'tri' | 'quad' | 'tetra' | 'hexa'; cellField: a Field holding the scalar of every mesh's cells."""
    with open(fileName, "w") as f:
        f.write("Title = \"%s\" \n" % title)
        f.write("variables = \"x\", \"y\", \"z\", \"%s\", \"cellCentroidY\" \n" % varName)
        for n, mesh in enumerate(meshes):
            raw = mesh.raw
            coords = np.asarray(raw.nodes).reshape(-1, 3)
            ncell, nnode = mesh.getCells().getSelfCount(), len(coords)
            f.write("Zone T = \"%s\" N = %s E = %s DATAPACKING = BLOCK, VARLOCATION = ([4-5]=CELLCENTERED), ZONETYPE=%s\n" %
                    ("nmesh%s" % n, nnode, ncell, TECPLOT_ZONE_TYPE[mtype]))

            def block(values):
                for i, v in enumerate(values):
                    f.write(_py2_str(v) + "    ")
                    if i % 5 == 4:
                        f.write("\n")
                f.write("\n")

            for d in range(3):
                block(coords[:, d])
            block(np.asarray(cellField[mesh.getCells()])[:ncell])
            block(np.asarray(geomFields.coordinate[mesh.getCells()])[:ncell, 1])
            row, col = cell_nodes(raw)
            for i in range(ncell):
                for v in col[row[i]:row[i + 1]]:
                    f.write(str(int(v) + 1) + "     ")
                f.write("\n")
            f.write("\n")
