"""The five accessors of the embedding-similarity path plus sem_partition_by, registered on pandas DataFrames under
the reference's names. Importing this package registers them."""
from ._common import register_all
from .load_sem_index import LoadSemIndexDataframe
from .sem_cluster_by import SemClusterByDataframe
from .sem_dedup import SemDedupByDataframe
from .sem_index import SemIndexDataframe
from .sem_partition_by import SemPartitionByDataframe
from .sem_search import SemSearchDataframe
from .sem_sim_join import SemSimJoinDataframe

__all__ = ["register_all", "LoadSemIndexDataframe", "SemClusterByDataframe", "SemDedupByDataframe", "SemIndexDataframe",
           "SemPartitionByDataframe", "SemSearchDataframe", "SemSimJoinDataframe"]
