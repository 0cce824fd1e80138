from .ffnn_model import FFNNModel
from .cnn_models import MinAtarModel, NatureAtariCnn
from .iqn_model import IQNModel
from .q_ensemble import QEnsemble
from .composite_model import CompositeModel
