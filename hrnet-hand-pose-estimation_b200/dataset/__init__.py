"""Device-side mirrors of the data-path glue the training / evaluation loops run around the network."""
