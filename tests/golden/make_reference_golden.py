"""Extract small golden fixtures from the reference's shipped lookup tables.

Run in the build container (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_reference_golden.py

Writes tests/golden/reference_tables.npz with selected rows of
  /root/reference/docs/data/Collagen12_BatchA.npy                       (recipe docs/data/README.md:20-42)
  /root/reference/docs/data/linear_reference_table/linear_reference_k2364.npy   (simulation.py:707-711)
These are the only result-pinning artefacts the reference holds for this path;
they were computed on a mesh that is not shipped, so they pin to mesh accuracy (few %) only.
The .npy files are pickles of plain numpy arrays (inspected with pickletools).
"""
import os

import numpy as np

REF = "/root/reference/docs/data"
HERE = os.path.dirname(os.path.abspath(__file__))
ROWS = [0, 10, 20, 30, 45, 60, 75, 90, 105, 120, 135, 149]

if __name__ == "__main__":
    a = np.load(os.path.join(REF, "Collagen12_BatchA.npy"), allow_pickle=True).item()
    lin = np.load(os.path.join(REF, "linear_reference_table", "linear_reference_k2364.npy"), allow_pickle=True).item()
    assert np.allclose(a["pressure"], np.logspace(-1, 4, 150), rtol=1e-14)
    np.savez_compressed(os.path.join(HERE, "reference_tables.npz"),
                        rows=np.array(ROWS), pressure=a["pressure"], x=a["x"],
                        batchA_y=a["y"][ROWS], batchA_material=np.array([1449.0, 0.00215, 0.032, 0.055]),
                        linear_pressure=lin["pressure"], linear_x=lin["x"], linear_y=lin["y"][ROWS],
                        linear_material=np.array([2364.0, 1e30, 1e30, 1e30]))
    print("wrote reference_tables.npz", a["y"].shape, lin["y"].shape)
