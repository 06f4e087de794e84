"""stark-backend_b200 — host-side mirror of the reference prover-backend interface over the
sm_100a C-ABI library (``libswirl_b200.so``).  The product path is the CUDA library; there is no
CPU fallback: loading fails loudly when the library is missing, and every compute call needs a
CUDA device."""
from .lib import load_library, library_path, SwirlError  # noqa: F401
from .backend import (  # noqa: F401
    AirProvingContext,
    B200Device,
    Transcript,
    WhirConfig,
    DeviceMatrix,
    PcsParams,
    StackedLayout,
    StackedPcsData,
    TraceTransporter,
    P,
    to_mont,
    from_mont,
)
from .prover import AirProvingKey, CommittedTraceData, Coordinator, Proof, SystemParams  # noqa: F401,E402
from . import codec  # noqa: F401,E402
from . import memory  # noqa: F401,E402
