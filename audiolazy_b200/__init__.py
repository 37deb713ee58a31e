"""audiolazy_b200 -- B200-native implementation of AudioLazy's linear-filter hot path.

Same names as the reference (``from audiolazy import ...``) for everything on the path:
``Stream``, ``thub``, ``Poly``, ``ZFilter``, ``z``, ``LinearFilter``, ``CascadeFilter``,
``ParallelFilter``, ``comb``, ``resonator``, ``lowpass``, ``highpass``, ``erb``,
``gammatone``, ``gammatone_erb_constants``, ``sHz``, ``almost_eq``; plus the bank object
the reference lacks (``FilterBank``, ``gammatone_bank``, ``erb_space``).

The per-sample recurrences run in hand-written sm_100a CUDA kernels behind the C ABI of
``include/alz_b200.h``; importing this package does not need a GPU, calling a filter does.
"""
from .core import StrategyDict
from .stream import Stream, StreamTeeHub, thub, tostream, avoid_stream
from .misc import sHz, almost_eq, zero_pad, elementwise, DEFAULT_SAMPLE_RATE
from .poly import Poly, x
from .filters import (LinearFilterProperties, LinearFilter, ZFilter, z, FilterList, CascadeFilter, ParallelFilter,
                      comb, resonator, lowpass, highpass)
from .auditory import erb, gammatone, gammatone_erb_constants, erb_space, gammatone_bank
from .bank import FilterBank, BankState
from .callers import envelope, maverage, karplus_strong, accumulate_z, zeros, ones, impulse, white_noise
from .io import chunks, WavStream, wav_batch, pcm_to_float32
from .linear_prediction import (ParCorError, acorr, lag_matrix, toeplitz, levinson_durbin, lpc, parcor, parcor_stable,
                                lsf, lsf_stable)

__version__ = "0.1.0"
