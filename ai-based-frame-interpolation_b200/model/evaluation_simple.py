"""Same metric surface as the reference's model/evaluation_simple.py:103-109 (it duplicates evaluation.py's)."""
try:
    from .evaluation import compute_metrics, compute_psnr, compute_ssim, evaluate_triplets  # noqa: F401
except ImportError:
    from evaluation import compute_metrics, compute_psnr, compute_ssim, evaluate_triplets  # noqa: F401
