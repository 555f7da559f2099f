"""Host-side mirror of /root/reference/metrabs_pytorch/multiperson/ for the steps either side of the crop model
(SURVEY.md 8f): crop generation, test-time-augmentation merge, plausibility filter + pose NMS.  All arithmetic on image
or pose data runs in libmetrabs_b200.so; torch supplies device memory and the tiny per-call parameter tensors."""
from metrabs_b200.multiperson.multiperson_model import Pose3dEstimator  # noqa: F401
