"""Import-only stub (no arithmetic) for torch-timeseries==0.1.10."""
