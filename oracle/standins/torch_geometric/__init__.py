"""Minimal stand-in for torch_geometric 2.6.1 (pinned by the reference's requirements.txt:26).

TEST INFRASTRUCTURE ONLY.  It exists so that the reference's three hot-path packages can be
imported verbatim from /root/reference in the authoring container (torch_geometric is not
installed and there is no network).  Semantics restated from the published PyG behaviour:
  propagate(edge_index, x, edge_attr): out[i] = sum_{e: edge_index[1,e]==i} message(x[edge_index[0,e]], edge_attr[e])
  global_add_pool / global_max_pool: segment sum / amax over `batch`, B = batch.max()+1
"""
from . import nn, utils  # noqa: F401
