"""Drop-in for the reference's ``lib/networks/enerf/network_human.py`` (ZJU-MoCap / interactive GUI
variant): ``network_module /root/repo/enerf_b200/network_human``.  Same model as
``enerf_b200.network.Network``; at the last cascade level only the rays inside
``batch['mask_at_box']`` are rendered and rgb is scattered back into a zero image
(network_human.py:90-107) -- done on device by csrc/mask_rays.cu."""
import os
import sys

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from enerf_b200.network import Network as _Base  # noqa: E402


class Network(_Base):
    def __init__(self):
        super().__init__()
        self.masked = True
