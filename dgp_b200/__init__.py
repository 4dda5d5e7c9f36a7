"""dgp_b200 -- B200-native (sm_100a CUDA) implementation of dgpsi's stochastic-imputation hot path behind the
reference's own API:  dgp(X, Y, all_layer).train(N);  emulator(all_layer, N).predict(x);  lgp / container / gp.
See DESIGN.md for the path, its boundary (include/dgpb.h) and what is out of scope."""
from .dgp import dgp
from .emulation import emulator
from .gp import gp
from .kernel_class import combine, kernel
from .likelihood_class import ZINB, ZIP, Categorical, Hetero, NegBin, Poisson
from .linkgp import container, lgp
from .utils import get_thread, nb_seed, read, set_thread, summary, write

__all__ = ["dgp", "gp", "emulator", "kernel", "combine", "container", "lgp", "write", "read", "summary", "nb_seed",
           "set_thread", "get_thread", "Poisson", "Hetero", "NegBin", "Categorical", "ZIP", "ZINB"]
__version__ = "0.1.0"
