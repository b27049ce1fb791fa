"""fcd_b200 -- B200-native (sm_100a) hot path of mehdirabiee/fcd behind the reference's own API:
get_model(params), CombinedLoss(params, device), sliding_window_inference(...)."""
from .config import get_default_params  # noqa: F401
from .evaluation import evaluate, evaluate_subject  # noqa: F401
from .get_loss import CombinedLoss, get_loss_function_from_params  # noqa: F401
from .get_model import get_model  # noqa: F401
from .inferers import post_process, post_process_segment, sliding_window_inference  # noqa: F401
from ._lib import PipelineError, check_errors  # noqa: F401
from .metrics import calculate_voxel_level_metrics, evaluate_fp  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from .sampling import GpuPatchSampler  # noqa: F401
