"""Builds tests/golden/kat_clusters.npz: the inputs of the reference's only known-answer test for this path
(/root/reference/tests/tangram_test.py:67-103 -- map_cells_to_space, mode='clusters', cluster_label=
'subclass_label', random_state=42, 500 epochs on data/test_ad_sc.h5ad x data/test_ad_sp.h5ad) together with the
expected `ad_map.X[0,0]` values the reference asserts.  Run in the build container only (needs /root/reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tangram_b200 as tg  # noqa: E402
from tangram_b200.h5ad import read_h5ad  # noqa: E402

DATA = "/root/reference/data"
# (lambda_g1, lambda_g2, lambda_d, density_prior, scale, expected X[0,0])  tests/tangram_test.py:70-78
CASES = [
    (1, 0, 0, None, True, 8.280743e-06),
    (1, 0, 0, None, False, 2.785552e-07),
    (1, 1, 0, None, True, 8.376801e-06),
    (1, 1, 0, None, False, 2.4095453e-07),
    (1, 1, 1, "uniform", True, 8.376801e-06),
    (1, 1, 1, "uniform", False, 2.4095453e-07),
    (1, 0, 2, "uniform", True, 1.3842443e-06),
    (1, 0, 1, "rna_count_based", True, 0.0023217443),
    (1, 0, 1, "uniform", True, 8.280743e-06),
]


def main():
    ad_sc = read_h5ad(os.path.join(DATA, "test_ad_sc.h5ad"))
    ad_sp = read_h5ad(os.path.join(DATA, "test_ad_sp.h5ad"))
    genes = ad_sc.uns["training_genes"]
    assert list(genes) == list(ad_sp.uns["training_genes"])
    out = dict(G=np.asarray(ad_sp[:, genes].X, dtype=np.float32),
               rna_count_based_density=np.asarray(ad_sp.obs["rna_count_based_density"], dtype=np.float64),
               uniform_density=np.asarray(ad_sp.obs["uniform_density"], dtype=np.float64))
    for scale in (True, False):
        agg = tg.adata_to_cluster_expression(ad_sc, "subclass_label", scale=scale, add_density=True)
        out["S_scale" if scale else "S_mean"] = np.asarray(agg[:, genes].X, dtype=np.float32)
        out["cluster_density"] = np.asarray(agg.obs["cluster_density"], dtype=np.float64)
        out["labels"] = np.array([str(x) for x in agg.obs["subclass_label"]])
    out["cases"] = np.array([(a, b, c, {None: 0, "uniform": 1, "rna_count_based": 2}[d], int(s), e)
                             for a, b, c, d, s, e in CASES], dtype=np.float64)
    path = os.path.join(HERE, "kat_clusters.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", out["S_scale"].shape, out["G"].shape, out["labels"][:4])


if __name__ == "__main__":
    main()
