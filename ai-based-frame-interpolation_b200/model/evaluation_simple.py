"""The reference's model/evaluation_simple.py duplicates evaluation.py's metric functions, baselines and evaluation
loop without the plotting half; here it is the same module under the reference's second set of names."""
try:
    from .evaluation import (METHODS, compute_metrics, compute_psnr, compute_ssim, evaluate_model,  # noqa: F401
                             evaluate_model_simple, evaluate_triplets, linear_interpolation_baseline, load_test_triplets,
                             main, optical_flow_interpolation_baseline, print_simple_summary, save_simple_results)
except ImportError:
    from evaluation import (METHODS, compute_metrics, compute_psnr, compute_ssim, evaluate_model,  # noqa: F401
                            evaluate_model_simple, evaluate_triplets, linear_interpolation_baseline, load_test_triplets,
                            main, optical_flow_interpolation_baseline, print_simple_summary, save_simple_results)

if __name__ == "__main__":
    raise SystemExit(main())
