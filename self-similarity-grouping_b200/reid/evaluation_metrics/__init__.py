from .ranking import cmc, mean_ap  # noqa: F401

__all__ = ['cmc', 'mean_ap']
