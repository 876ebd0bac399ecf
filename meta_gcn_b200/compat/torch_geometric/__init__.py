"""PyG-1.3-shaped operator namespace on libmgcn (see meta_gcn_b200.compat)."""
__version__ = "1.3.2+mgcn"
