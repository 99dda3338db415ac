"""llamole_b200 -- B200-native (sm_100a) graph-module hot path of Llamole.

    from llamole_b200 import GraphDiT, GraphCLIP, GraphPredictor

are drop-ins for the reference classes of the same names under src/model/graph_{decoder,encoder,predictor}.
"""
from .graph_decoder import GraphDiT, set_smiles_backend  # noqa: F401
from .graph_encoder import GraphCLIP  # noqa: F401
from .graph_predictor import GraphPredictor  # noqa: F401
from .condition_queue import ConditionQueue  # noqa: F401

__version__ = "0.1.0"
