{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE EmptyDataDecls #-}
-- | GlomeTrace.CUDA, bottom layer: the flat-scene builder handle and one emitter per C constructor of
-- include/glome_cuda.h (`glome_sb_*`).  This module sits BELOW Data.Glome.Solid in the import graph (it knows
-- nothing about `Solid`), so that `Solid` can carry the `flatten` method (haskell/glometrace-cuda.patch) and every
-- primitive module can implement it with the emitters below.
--
-- NOT COMPILED IN THIS REPOSITORY'S CI (no GHC in the build image, SURVEY.md F2).  The C side of every import is
-- exercised by tests/test_host_builder.py through ctypes with the same argument layouts.
module Data.Glome.CUDA.Flat
  ( GlomeBuilderH, FlatBuilder(..), TexDesc(..)
  , chk
  , emitVoid, emitSphere, emitTriangle, emitTriangleNorm, emitBox, emitPlane, emitDisc, emitCylinder, emitCone
  , emitList, emitInstance, emitDifference, emitIntersection, emitTex, emitTag, emitNoShadow, emitOnlyShadow
  , emitBound, emitInnerBound, emitBihPrebuilt, emitMeshPrebuilt
  , c_builder_create, c_builder_destroy, c_builder_set_build_device, c_sb_flatten
  , c_sb_mat_surface, c_sb_mat_reflect, c_sb_mat_refract, c_sb_mat_warp, c_sb_mat_warp_set_scene, c_sb_mat_additive
  , c_sb_mat_blend, c_sb_tex_uniform, c_sb_tex_stripe_blend, c_sb_tex_perlin_blend, c_sb_light, c_sb_lightset
  , c_last_error, vec3, xfm24, bbox6
  ) where

import Control.Exception (throwIO, ErrorCall(..))
import Data.Int
import Foreign
import Foreign.C.String
import Foreign.C.Types

import Data.Glome.Vec

data GlomeBuilderH   -- opaque GlomeBuilder (include/glome_cuda.h)

-- | A texture the device can evaluate (the closures of `Texture t m = Ray -> Rayint t m -> m`, Solid.hs:97, cannot be
-- looked into).  The vocabulary is the one glome_cuda.h reifies: GLOME_TEX_UNIFORM, _STRIPE_BLEND, _PERLIN_BLEND.
data TexDesc m
  = TexUniform m               -- ^ @\\_ _ -> m@                                          (Shader.hs:55 t_uniform)
  | TexStripeBlend m m Vec     -- ^ @Blend a b (triangle_wave (vdot pos axis))@           (TestScene.hs:225 t_stripe)
  | TexPerlinBlend m m Flt     -- ^ @Blend a b (perlin (vscale pos s))@                   (TestScene.hs:214 t_mottled)

-- | What `flatten` threads through the scene graph.  @tex@ is @Texture t m@, @tag@ is @t@, @mat@ is @m@.
data FlatBuilder tex tag mat = FlatBuilder
  { fbHandle  :: Ptr GlomeBuilderH
  , fbTexture :: tex -> IO Int          -- ^ texture closure -> device texture id (Data.Glome.CUDA supplies a registry)
  , fbTexDesc :: TexDesc mat -> IO Int  -- ^ described texture -> device texture id (materials are reified on the way)
  , fbTag     :: tag -> IO Int          -- ^ tag value -> dense device tag id
  }

foreign import ccall unsafe "glome_last_error"        c_last_error        :: IO CString
foreign import ccall unsafe "glome_builder_create"    c_builder_create    :: Ptr (Ptr GlomeBuilderH) -> IO CInt
foreign import ccall unsafe "glome_builder_destroy"   c_builder_destroy   :: Ptr GlomeBuilderH -> IO CInt
foreign import ccall unsafe "glome_builder_set_build_device" c_builder_set_build_device :: Ptr GlomeBuilderH -> CInt -> IO CInt
foreign import ccall safe   "glome_sb_flatten"        c_sb_flatten        :: Ptr GlomeBuilderH -> CInt -> Ptr () -> IO CInt

foreign import ccall unsafe "glome_sb_void"           c_sb_void           :: Ptr GlomeBuilderH -> IO CInt
foreign import ccall unsafe "glome_sb_sphere"         c_sb_sphere         :: Ptr GlomeBuilderH -> Ptr CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_triangle"       c_sb_triangle       :: Ptr GlomeBuilderH -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_trianglenorm"   c_sb_trianglenorm   :: Ptr GlomeBuilderH -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_box"            c_sb_box            :: Ptr GlomeBuilderH -> Ptr CDouble -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_plane_offset"   c_sb_plane_offset   :: Ptr GlomeBuilderH -> Ptr CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_disc_raw"       c_sb_disc_raw       :: Ptr GlomeBuilderH -> Ptr CDouble -> Ptr CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_cylinder_z"     c_sb_cylinder_z     :: Ptr GlomeBuilderH -> CDouble -> CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_cone_z"         c_sb_cone_z         :: Ptr GlomeBuilderH -> CDouble -> CDouble -> CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_list"           c_sb_list           :: Ptr GlomeBuilderH -> CInt -> Ptr Int32 -> IO CInt
foreign import ccall unsafe "glome_sb_instance"       c_sb_instance       :: Ptr GlomeBuilderH -> CInt -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_difference_ex"  c_sb_difference_ex  :: Ptr GlomeBuilderH -> CInt -> CInt -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_intersection"   c_sb_intersection   :: Ptr GlomeBuilderH -> CInt -> Ptr Int32 -> IO CInt
foreign import ccall unsafe "glome_sb_tex"            c_sb_tex            :: Ptr GlomeBuilderH -> CInt -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_tag"            c_sb_tag            :: Ptr GlomeBuilderH -> CInt -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_noshadow"       c_sb_noshadow       :: Ptr GlomeBuilderH -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_onlyshadow"     c_sb_onlyshadow     :: Ptr GlomeBuilderH -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_bound_object"   c_sb_bound_object   :: Ptr GlomeBuilderH -> CInt -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_innerbound"     c_sb_innerbound     :: Ptr GlomeBuilderH -> CInt -> CInt -> IO CInt
-- trees the Haskell constructors already built, as pre-order streams (glome_cuda.h: glome_sb_bih_prebuilt)
foreign import ccall safe   "glome_sb_bih_prebuilt"   c_sb_bih_prebuilt   :: Ptr GlomeBuilderH -> Int64 -> Ptr Int32 -> Int64 -> Ptr Int32 -> Ptr CDouble -> Ptr CDouble -> IO CInt
foreign import ccall safe   "glome_sb_mesh_prebuilt"  c_sb_mesh_prebuilt  :: Ptr GlomeBuilderH -> Int64 -> Ptr CDouble -> Int64 -> Ptr CDouble -> Int64 -> Ptr Int32 -> CInt -> Ptr Int32 -> CInt -> Ptr Int32 -> Int64 -> Ptr Int32 -> Ptr CDouble -> Int64 -> Ptr Int32 -> Ptr CDouble -> IO CInt
-- materials, textures, lights
foreign import ccall unsafe "glome_sb_mat_surface"    c_sb_mat_surface    :: Ptr GlomeBuilderH -> Ptr CDouble -> CDouble -> CDouble -> CDouble -> CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_mat_reflect"    c_sb_mat_reflect    :: Ptr GlomeBuilderH -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_mat_refract"    c_sb_mat_refract    :: Ptr GlomeBuilderH -> CDouble -> CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_mat_warp"       c_sb_mat_warp       :: Ptr GlomeBuilderH -> CInt -> CInt -> CInt -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_mat_warp_set_scene" c_sb_mat_warp_set_scene :: Ptr GlomeBuilderH -> CInt -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_mat_additive"   c_sb_mat_additive   :: Ptr GlomeBuilderH -> CInt -> Ptr Int32 -> IO CInt
foreign import ccall unsafe "glome_sb_mat_blend"      c_sb_mat_blend      :: Ptr GlomeBuilderH -> CInt -> CInt -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_tex_uniform"    c_sb_tex_uniform    :: Ptr GlomeBuilderH -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_tex_stripe_blend" c_sb_tex_stripe_blend :: Ptr GlomeBuilderH -> CInt -> CInt -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_tex_perlin_blend" c_sb_tex_perlin_blend :: Ptr GlomeBuilderH -> CInt -> CInt -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_light"          c_sb_light          :: Ptr GlomeBuilderH -> Ptr CDouble -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_lightset"       c_sb_lightset       :: Ptr GlomeBuilderH -> CInt -> Ptr Int32 -> IO CInt

-- | Negative return codes carry a message (GLOME_EBUILD mirrors the reference's `error` calls).
chk :: String -> CInt -> IO Int
chk what rc
  | rc >= 0   = return (fromIntegral rc)
  | otherwise = do msg <- c_last_error >>= peekCString
                   throwIO (ErrorCall ("GlomeTrace.CUDA: " ++ what ++ ": " ++ msg))

vec3 :: Vec -> [CDouble]
vec3 (Vec x y z) = map realToFrac [x, y, z]

-- | @Xfm fwd inv@ (Vec.hs:407-414) as the 24 doubles of glome_cuda.h: the forward 3x4 matrix row by row, then the inverse.
xfm24 :: Xfm -> [CDouble]
xfm24 (Xfm f i) = mat f ++ mat i
  where mat (Matrix a b c d  e f' g h  i' j k l) = map realToFrac [a, b, c, d, e, f', g, h, i', j, k, l]

bbox6 :: Bbox -> [CDouble]
bbox6 (Bbox a b) = vec3 a ++ vec3 b

ci :: Int -> CInt
ci = fromIntegral

ids32 :: [Int] -> [Int32]
ids32 = map fromIntegral

emitVoid :: FlatBuilder a b c -> IO Int
emitVoid fb = c_sb_void (fbHandle fb) >>= chk "Void"

-- | Sphere c r (1/r): the device recomputes nothing it does not need; 1/r is not stored there (Sphere.hs:11).
emitSphere :: FlatBuilder a b c -> Vec -> Flt -> IO Int
emitSphere fb c r = withArray (vec3 c) $ \p -> c_sb_sphere (fbHandle fb) p (realToFrac r) >>= chk "Sphere"

emitTriangle :: FlatBuilder a b c -> Vec -> Vec -> Vec -> IO Int
emitTriangle fb p1 p2 p3 = withArray (concatMap vec3 [p1, p2, p3]) $ \p -> c_sb_triangle (fbHandle fb) p >>= chk "Triangle"

emitTriangleNorm :: FlatBuilder a b c -> Vec -> Vec -> Vec -> Vec -> Vec -> Vec -> IO Int
emitTriangleNorm fb p1 p2 p3 n1 n2 n3 =
  withArray (concatMap vec3 [p1, p2, p3, n1, n2, n3]) $ \p -> c_sb_trianglenorm (fbHandle fb) p >>= chk "TriangleNorm"

-- | Box (Bbox p1 p2): the corners are already ordered (Box.hs:12-15); fmin / fmax of ordered corners is the identity.
emitBox :: FlatBuilder a b c -> Bbox -> IO Int
emitBox fb (Bbox a b) = withArray (vec3 a) $ \pa -> withArray (vec3 b) $ \pb -> c_sb_box (fbHandle fb) pa pb >>= chk "Box"

-- | Plane norm offset as stored (Plane.hs:11): `plane_offset`, no renormalisation.
emitPlane :: FlatBuilder a b c -> Vec -> Flt -> IO Int
emitPlane fb n off = withArray (vec3 n) $ \p -> c_sb_plane_offset (fbHandle fb) p (realToFrac off) >>= chk "Plane"

-- | Disc pos norm (r*r) as stored (Cone.hs:21).
emitDisc :: FlatBuilder a b c -> Vec -> Vec -> Flt -> IO Int
emitDisc fb pos n rsqr =
  withArray (vec3 pos) $ \pp -> withArray (vec3 n) $ \pn -> c_sb_disc_raw (fbHandle fb) pp pn (realToFrac rsqr) >>= chk "Disc"

emitCylinder :: FlatBuilder a b c -> Flt -> Flt -> Flt -> IO Int
emitCylinder fb r h1 h2 = c_sb_cylinder_z (fbHandle fb) (realToFrac r) (realToFrac h1) (realToFrac h2) >>= chk "Cylinder"

emitCone :: FlatBuilder a b c -> Flt -> Flt -> Flt -> Flt -> IO Int
emitCone fb r c1 c2 h = c_sb_cone_z (fbHandle fb) (realToFrac r) (realToFrac c1) (realToFrac c2) (realToFrac h) >>= chk "Cone"

-- | A list of solids as it is (Solid.hs:326): no `group` flattening, Voids kept.
emitList :: FlatBuilder a b c -> [Int] -> IO Int
emitList fb xs = withArrayLen (ids32 xs) $ \n p -> c_sb_list (fbHandle fb) (ci n) p >>= chk "[s]"

emitInstance :: FlatBuilder a b c -> Int -> Xfm -> IO Int
emitInstance fb s x = withArray (xfm24 x) $ \p -> c_sb_instance (fbHandle fb) (ci s) p >>= chk "Instance"

emitDifference :: FlatBuilder a b c -> Int -> Int -> Bool -> IO Int
emitDifference fb a b useatex = c_sb_difference_ex (fbHandle fb) (ci a) (ci b) (if useatex then 1 else 0) >>= chk "Difference"

emitIntersection :: FlatBuilder a b c -> [Int] -> IO Int
emitIntersection fb xs = withArrayLen (ids32 xs) $ \n p -> c_sb_intersection (fbHandle fb) (ci n) p >>= chk "Intersection"

emitTex, emitTag :: FlatBuilder a b c -> Int -> Int -> IO Int
emitTex fb s t = c_sb_tex (fbHandle fb) (ci s) (ci t) >>= chk "Tex"
emitTag fb s t = c_sb_tag (fbHandle fb) (ci s) (ci t) >>= chk "Tag"

emitNoShadow, emitOnlyShadow :: FlatBuilder a b c -> Int -> IO Int
emitNoShadow fb s = c_sb_noshadow (fbHandle fb) (ci s) >>= chk "NoShadow"
emitOnlyShadow fb s = c_sb_onlyshadow (fbHandle fb) (ci s) >>= chk "OnlyShadow"

emitBound, emitInnerBound :: FlatBuilder a b c -> Int -> Int -> IO Int
emitBound fb a b = c_sb_bound_object (fbHandle fb) (ci a) (ci b) >>= chk "Bound"
emitInnerBound fb a b = c_sb_innerbound (fbHandle fb) (ci a) (ci b) >>= chk "InnerBound"

-- | `Bih bb root` with the tree the constructor built (Bih.hs:51-57, 309-324).  @kinds@ / @splits@: one record per
-- tree node in pre-order -- a branch is (axis, (lsplit, rsplit)), a leaf of n items is (-(n+1), (0, 0)); @items@: the
-- flattened leaf items in the order the leaves hold them.
emitBihPrebuilt :: FlatBuilder a b c -> [Int] -> [Int] -> [(Flt, Flt)] -> Bbox -> IO Int
emitBihPrebuilt fb items kinds splits bb =
  withArrayLen (ids32 items) $ \ni pitems ->
  withArrayLen (ids32 kinds) $ \nk pkinds ->
  withArray (concat [ [realToFrac l, realToFrac r] | (l, r) <- splits ]) $ \psplits ->
  withArray (bbox6 bb) $ \pbb ->
    c_sb_bih_prebuilt (fbHandle fb) (fromIntegral ni) pitems (fromIntegral nk) pkinds psplits pbb >>= chk "Bih"

-- | `Mesh verts norms tris texs tags bb bvh` with the BVH the constructor built (Mesh.hs:36-42).  @tris@: 8 ints per
-- `Tri a b c n1 n2 n3 tex tag`; @kinds@ / @boxes@: one record per BVH node in pre-order -- a branch is (0, lbb ++ rbb as
-- 12 doubles), a leaf of n triangles is (-(n+1), 12 zeros); @leafTris@: the leaves' triangle indices in that order.
emitMeshPrebuilt :: FlatBuilder a b c -> [Vec] -> [Vec] -> [[Int]] -> [Int] -> [Int] -> [Int] -> [[CDouble]] -> [Int] -> Bbox -> IO Int
emitMeshPrebuilt fb verts norms tris texs tags kinds boxes leafTris bb =
  withArrayLen (concatMap vec3 verts) $ \nv3 pverts ->
  withArrayLen (concatMap vec3 norms) $ \nn3 pnorms ->
  withArrayLen (ids32 (concat tris)) $ \nt8 ptris ->
  withArrayLen (ids32 texs) $ \ntex ptexs ->
  withArrayLen (ids32 tags) $ \ntag ptags ->
  withArrayLen (ids32 kinds) $ \nk pkinds ->
  withArray (concat boxes) $ \pboxes ->
  withArrayLen (ids32 leafTris) $ \nl pleaf ->
  withArray (bbox6 bb) $ \pbb ->
    c_sb_mesh_prebuilt (fbHandle fb) (fromIntegral (nv3 `div` 3)) pverts (fromIntegral (nn3 `div` 3)) pnorms
                       (fromIntegral (nt8 `div` 8)) ptris (ci ntex) ptexs (ci ntag) ptags
                       (fromIntegral nk) pkinds pboxes (fromIntegral nl) pleaf pbb >>= chk "Mesh"
