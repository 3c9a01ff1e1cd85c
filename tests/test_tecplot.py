"""Formats at the path's edge: the cell -> node connectivity in the elements' canonical node order
(Mesh::getCellNodes) against the reference's own, on generated meshes and on the reference's case files, and the
Tecplot finite-element file its parallel test scripts write (dumpTecplotFile, `temp_procN.dat`)."""
import os

import numpy as np
import pytest

from fvm_b200 import exporters as E, importers, meshgen as G, models as M

REF_TEST = "/root/reference/src/fvm/test"


def _ref_cell_nodes(ref, raw):
    rm = ref.RefMesh.from_raw(raw.dim, raw.n_cells, raw.nodes, raw.face_cells, raw.face_nodes, raw.face_node_count,
                              raw.face_group_size)
    return rm.cell_nodes()


@pytest.mark.parametrize("name", ["quad", "hex", "tet"])
def test_cell_nodes_follow_the_reference_element_order(ref, name):
    raw = {"quad": lambda: G.quad_mesh(9, 7, jitter=0.15, seed=1), "hex": lambda: G.hex_mesh(5, 4, 6, jitter=0.15, seed=2),
           "tet": lambda: G.tet_mesh(4, 3, 5)}[name]()
    rr, rc = _ref_cell_nodes(ref, raw)
    r, c = E.cell_nodes(raw)
    assert np.array_equal(r, rr) and np.array_equal(c, rc)
    assert len(set(np.diff(r))) == 1 and np.diff(r)[0] == {"quad": 4, "hex": 8, "tet": 4}[name]


@pytest.mark.parametrize("cas", ["cav32.cas", "tri_894.cas", "cav_tetra.cas", "cav_hexa.cas"])
def test_cell_nodes_of_the_reference_case_files(ref, cas):
    """quadrilaterals, triangles and tetrahedra of the reference's own case files: fvm_b200.importers numbers the nodes
    of the cell zone in the order its cells meet them, as the reference's reader does (I/FluentReader.cpp:841-856), so
    node coordinates and cell -> node lists are identical arrays."""
    path = os.path.join(REF_TEST, cas)
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    fc = importers.FluentCase(path)
    fc.read()
    raw = fc.getMeshList()[0].raw
    rm = ref.RefMesh.from_cas(path)
    rr, rc = rm.cell_nodes()
    r, c = E.cell_nodes(raw)
    assert np.array_equal(r, rr) and np.array_equal(c, rc)
    assert np.array_equal(np.asarray(raw.nodes).reshape(-1, 3), rm.node_coordinates())


def test_tecplot_file_layout(hostsim_lib, tmp_path):
    raw = G.quad_mesh(6, 4, jitter=0.1, seed=3)
    mesh = M.Mesh(raw)
    geom = M.GeomFields("geom")
    M.MeshMetricsCalculatorA(geom, [mesh], lib=hostsim_lib).init()
    tf = M.ThermalFields("therm")
    cells = mesh.getCells()
    tf.temperature[cells] = 300.0 + np.arange(cells.getCount()) / 7.0
    out = tmp_path / "temp_proc0.dat"
    E.dumpTecplotFile(str(out), [mesh], "quad", tf.temperature, geom)
    lines = out.read_text().split("\n")
    assert lines[0] == 'Title = " tecplot file for 2D Cavity problem " '
    assert lines[1] == 'variables = "x", "y", "z", "velX", "cellCentroidY" '
    nn, nc = len(raw.nodes), raw.n_cells
    assert lines[2] == ('Zone T = "nmesh0" N = %d E = %d DATAPACKING = BLOCK, VARLOCATION = ([4-5]=CELLCENTERED), '
                        'ZONETYPE=FEQUADRILATERAL' % (nn, nc))
    body = out.read_text().split("\n", 3)[3].split()
    vals = body[:3 * nn + 2 * nc]
    x = np.array(vals[:nn], float)
    assert np.abs(x - np.asarray(raw.nodes)[:, 0]).max() <= 1e-11
    t = np.array(vals[3 * nn:3 * nn + nc], float)
    assert np.abs(t - (300.0 + np.arange(nc) / 7.0)).max() <= 1e-9
    assert all("." in v or "e" in v for v in vals)            # Python 2's str(float): never a bare integer
    conn = np.array(body[3 * nn + 2 * nc:], int).reshape(nc, 4)
    r, c = E.cell_nodes(raw)
    assert np.array_equal(conn, c.reshape(nc, 4) + 1)
    # five values per line in the data blocks, one cell per line in the connectivity
    assert len(lines[3].split()) == 5
    assert E._py2_str(1.0) == "1.0" and E._py2_str(0.1 + 0.2) == "0.3" and E._py2_str(1e-20) == "1e-20"
