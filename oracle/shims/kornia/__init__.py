"""Test-infrastructure shim: the reference imports ``kornia.utils.create_meshgrid`` only
(/root/reference/lib/networks/enerf/utils.py:4,65). kornia is not installed in this image."""
