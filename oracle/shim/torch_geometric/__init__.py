"""ORACLE / TEST INFRASTRUCTURE ONLY.  Minimal CPU restatement of the torch_geometric 1.3.x surface
that /root/reference imports.  Never imported by meta_gcn_b200/."""
__version__ = "1.3.2-oracle-shim"
