"""`project_genes` with the reference's contract (/root/reference/tangram/utils.py:338-374)."""
import numpy as np

from . import mapping_utils as mu
from .adata import make_adata


def project_genes(adata_map, adata_sc, cluster_label=None, scale=True):
    adata_sc.var.index = [g.lower() for g in adata_sc.var.index]                 # :353
    adata_sc.var_names_make_unique()                                              # :356
    keep = np.asarray((adata_sc.X != 0).sum(axis=0)).reshape(-1) >= 1             # :359
    if not keep.all():
        adata_sc = adata_sc[:, keep]
    if cluster_label:
        adata_sc = mu.adata_to_cluster_expression(adata_sc, cluster_label, scale=scale)
    if not adata_map.obs.index.equals(adata_sc.obs.index):
        raise ValueError("The two AnnDatas need to have same `obs` index.")
    X = adata_sc.X.toarray() if hasattr(adata_sc.X, "toarray") else np.asarray(adata_sc.X)
    mapper = getattr(adata_map, "_tgb200_mapper", None)
    if mapper is not None and mapper.n_cells == X.shape[0]:
        X_space = mapper.project(X)               # softmax(M)^T X on the device (:368 is a host GEMM)
    else:
        X_space = np.asarray(adata_map.X).T @ X
    adata_ge = make_adata(X=X_space, obs=adata_map.var, var=adata_sc.var, uns=adata_sc.uns)
    training_genes = adata_map.uns["train_genes_df"].index.values
    adata_ge.var["is_training"] = adata_ge.var.index.isin(training_genes)
    return adata_ge
