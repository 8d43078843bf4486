"""Mirror of the reference's `mtl` plugin surface (mtl.apis / mtl.data / mtl.engine /
mtl.runner / mtl.utils / mtl.model), re-implemented without mmcv."""
