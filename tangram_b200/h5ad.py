"""`read_h5ad` for the anndata-0.7 on-disk layout used by Tangram's fixtures and tutorials
(/root/reference/data/*.h5ad), on top of the pure-Python HDF5 reader.
Returns a MiniAnnData (or a real AnnData when `anndata` is importable)."""
import numpy as np
import pandas as pd

from .adata import MiniAnnData
from .h5mini import H5File


def _read_df(f, group):
    attrs = f.attrs(group)
    names = f.listdir(group)
    index_key = attrs.get("_index", "_index")
    order = attrs.get("column-order")
    cols = [c for c in (list(order) if order is not None else names) if c in names]
    index = [str(x) for x in f.read(f"{group}/{index_key}").tolist()]
    data = {}
    cats = f.listdir(f"{group}/__categories") if "__categories" in names else []
    for c in cols:
        vals = f.read(f"{group}/{c}")
        if c in cats:
            categories = [str(x) for x in f.read(f"{group}/__categories/{c}").tolist()]
            codes = np.asarray(vals, dtype=np.int64)
            data[c] = pd.Categorical.from_codes(codes, categories)
        elif vals.dtype == object:
            data[c] = [str(x) for x in vals.tolist()]
        else:
            data[c] = vals
    return pd.DataFrame(data, index=index)


def _read_x(f):
    if f.is_group(f.resolve("/X")):
        import scipy.sparse as sp
        attrs = f.attrs("/X")
        shape = tuple(int(x) for x in attrs["shape"])
        data, indices, indptr = f.read("/X/data"), f.read("/X/indices"), f.read("/X/indptr")
        enc = attrs.get("encoding-type", "csr_matrix")
        cls = sp.csr_matrix if enc == "csr_matrix" else sp.csc_matrix
        return cls((data, indices, indptr), shape=shape)
    return f.read("/X")


def read_h5ad(path):
    f = H5File(path)
    X = _read_x(f)
    obs = _read_df(f, "/obs")
    var = _read_df(f, "/var")
    uns = {}
    if "/uns" in f:
        for k in f.listdir("/uns"):
            try:
                v = f.read(f"/uns/{k}")
                uns[k] = [str(x) for x in v.tolist()] if v.dtype == object else v
            except Exception:  # noqa: BLE001  (nested groups etc. are not needed by the mapping path)
                pass
    return MiniAnnData(X=X, obs=obs, var=var, uns=uns)
