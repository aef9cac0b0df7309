"""Host-side calibration helpers of the reference's src/helpers (the data loaders need the network and are not part
of the simulation path; the CDS bootstrap is)."""
