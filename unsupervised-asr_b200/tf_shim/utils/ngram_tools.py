"""`ngram2kernel` / `read_ngram` with the reference's signatures (utils/tools.py:255-279, 365-374), re-exported
from the TF-free package so that `from utils.tools import ngram2kernel` can be pointed here."""
from eodm_b200.tools import load_vocab, ngram2kernel, read_ngram  # noqa: F401
