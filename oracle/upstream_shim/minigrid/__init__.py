"""Restated subset of the public `minigrid` API (TEST INFRASTRUCTURE ONLY).

See oracle/upstream_shim/README.md.  Semantics follow SURVEY.md Appendix A.
"""
__shim__ = True
