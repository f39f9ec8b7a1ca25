from .evaluate import evaluate, exact_evaluate, mean_logs

__all__ = ['evaluate', 'exact_evaluate', 'mean_logs']
