"""Model-side plugin surface of the reference (core/models): add_uncertainty wrapper, quantile head, UNet trunk."""
