"""lotus_b200 — B200-native (sm_100a) backend for the embedding-similarity hot path of lotus-data/lotus:
`lotus.vector_store.FaissVS` + `sem_index / sem_search / sem_sim_join / sem_dedup / sem_cluster_by`.

    import lotus_b200 as lotus                      # same names as the reference for this path
    lotus.settings.configure(rm=rm, vs=lotus.B200VS())
    df.sem_index("text", "idx_dir").sem_sim_join(other, "a", "b", K=32)

With the real `lotus` package importable, call `lotus_b200.install()` instead: it plugs `B200VS` into
`lotus.settings`, replaces `lotus.utils.cluster` and re-registers the accessors that need the streaming kernels.
The compute lives in libb2lotus.so (include/lotus_b200.h); there is no CPU fallback.
"""
from . import utils
from .rm import RM, BagOfWordsRM, HashRM, TableRM
from .settings import settings
from .types import RMOutput
from .vs import METRIC_INNER_PRODUCT, METRIC_L2, VS, B200VS
from . import sem_ops  # noqa: F401  (registers the accessors)

__version__ = "0.1.0"


def install(vs: "B200VS | None" = None, **vs_kwargs):
    """Make B200VS the vector store of whichever settings object the operators read (the real lotus.settings when
    lotus is importable, ours otherwise) and route lotus.utils.cluster to the device k-means."""
    store = vs if vs is not None else B200VS(**vs_kwargs)
    try:
        import lotus as _ref  # type: ignore
        _ref.settings.configure(vs=store)
        _ref.utils.cluster = utils.cluster
    except Exception:
        settings.configure(vs=store)
    return store


__all__ = ["RM", "BagOfWordsRM", "HashRM", "TableRM", "VS", "B200VS", "RMOutput", "settings", "utils", "install",
           "METRIC_INNER_PRODUCT", "METRIC_L2"]
