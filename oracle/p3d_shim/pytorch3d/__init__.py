"""Minimal CPU restatement of the PyTorch3D v0.6.2 API surface that korrawe/harp imports.  ORACLE / TEST ONLY."""
__version__ = "0.6.2+harp-oracle-shim"
