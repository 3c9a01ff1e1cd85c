"""Extract the reference partitioner's registered golden for cav32.cas on 4 ranks
(T/PARALLEL_TESTS/PARTHMESH/QUAD_1024/proc4/GOLDEN, test CAVITY_QUAD1024_PROCS4 of T/PARALLEL_TESTS/TESTS) into a small
fixture: the ParMETIS cell assignment (`_cellParts` of procK_debug_print.dat) and, per rank, the ghost-cell maps of
the partitioned mesh the reference built from it (mesh_procK_info.dat: "neightMeshID = p  <local ghost cell> ===>
<cell in p's local numbering>"). ParMETIS itself is a missing blob of the reference tree, so the assignment
can only be REPLAYED: fvm_b200.partition.partition_mesh must then produce exactly these maps
(tests/test_partition.py::test_partition_replays_the_reference_parthmesh_golden).

  python tests/golden/make_parthmesh_golden.py      (needs /root/reference)
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
D = "/root/reference/src/fvm/test/PARALLEL_TESTS/PARTHMESH/QUAD_1024/proc4/GOLDEN/"


def main():
    nparts = 4
    txt = open(D + "proc0_debug_print.dat").read()
    a, b = txt.index("_cellParts :"), txt.index("_faceParts :")
    rows = re.findall(r"row\[(\d+)\] = (\d+)\s+(\d+)", txt[a:b])
    part = np.array([int(r[2]) for r in rows], np.int32)
    assert [int(r[0]) for r in rows] == list(range(len(rows)))
    out = dict(cell_parts=part, nparts=nparts)
    for k in range(nparts):
        m = re.findall(r"neightMeshID = (\d+)\s+(\d+)\s+===>\s+(\d+)", open(D + "mesh_proc%d_info.dat" % k).read())
        out["halo%d" % k] = np.array([[int(x) for x in t] for t in m], np.int32)   # (peer, local ghost, peer's cell)
        head = open(D + "mesh_proc%d.dat" % k).read().splitlines()[2]
        out["nodes_cells%d" % k] = np.array([int(v) for v in re.findall(r"[NE] = (\d+)", head)], np.int32)
    np.savez_compressed(os.path.join(HERE, "parthmesh_quad1024_proc4.npz"), **out)
    print("parthmesh_quad1024_proc4.npz: %d cells + ghosts assigned, interface entries per rank %s"
          % (len(part), [len(out["halo%d" % k]) for k in range(nparts)]))


if __name__ == "__main__":
    main()
