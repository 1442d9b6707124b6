from trlda_b200.models import Distribution, LDA, OnlineLDA, BatchLDA, CumulativeLDA

__all__ = ['Distribution', 'LDA', 'OnlineLDA', 'BatchLDA', 'CumulativeLDA']
