for mb in 0 36 24 12; do echo "chunk $mb MB:"; RV3D_RASTER_CHUNK_MB=$mb python tools/dev_time_raster.py; done
