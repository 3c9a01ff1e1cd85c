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
