"""Drop-in functions for the reference's ``data_load_preprocess`` package."""
