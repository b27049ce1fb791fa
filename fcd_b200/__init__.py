"""fcd_b200 -- B200-native (sm_100a) hot path of mehdirabiee/fcd behind the reference's own API:
get_model(params), CombinedLoss(params, device), sliding_window_inference(...)."""
from .config import get_default_params  # noqa: F401
from .get_loss import CombinedLoss, get_loss_function_from_params  # noqa: F401
from .get_model import get_model  # noqa: F401
