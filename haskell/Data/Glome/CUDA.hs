{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE EmptyDataDecls #-}
-- | GlomeTrace.CUDA: the reference-side binding to libglomecuda.so (include/glome_cuda.h).
--
-- NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no GHC (SURVEY.md F2).  This module is
-- the binding a maintainer adds to the GlomeTrace package (see INTEGRATION.md); it only uses
-- `foreign import ccall` over the C shim, as BASELINE.json's north_star asks.
--
-- Scene construction stays in Haskell: build the scene with the usual GlomeTrace constructors
-- (sphere, box, cone, difference, bih, mesh, tex, transform ...), call 'flattenScene' once, then
-- render frames / trace ray batches on the B200.
module Data.Glome.CUDA
  ( CudaScene
  , FlatBuilder, Flatten(..)
  , withCudaScene, newCudaScene
  , renderTilesCuda, RenderMode(..), RenderOpts(..), defaultRenderOpts
  , rayintBatch, shadowBatch, traceBatch
  , CudaHit(..)
  ) where

import Control.Concurrent.MVar
import Control.Exception (throwIO, ErrorCall(..))
import Control.Monad (when)
import Data.Int
import Data.Word
import Foreign
import Foreign.C.String
import Foreign.C.Types

import Data.Glome.Vec
import Data.Glome.Scene (Camera(..))

-- ---------------------------------------------------------------------------------------------
-- C side (include/glome_cuda.h)
-- ---------------------------------------------------------------------------------------------
data GlomeSceneH     -- opaque GlomeScene
data GlomeMultiH
data GlomeBuilderH   -- opaque GlomeBuilder
data GlomeFlatSceneC -- struct GlomeFlatScene (filled by glome_sb_flatten)

-- calls block on the GPU: import them `safe` so other Haskell threads keep running
foreign import ccall safe   "glome_scene_create"   c_scene_create   :: Ptr GlomeFlatSceneC -> CInt -> Ptr (Ptr GlomeSceneH) -> IO CInt
foreign import ccall safe   "glome_scene_destroy"  c_scene_destroy  :: Ptr GlomeSceneH -> IO CInt
foreign import ccall safe   "&glome_scene_destroy" p_scene_destroy  :: FunPtr (Ptr GlomeSceneH -> IO ())
foreign import ccall safe   "glome_render"         c_render         :: Ptr GlomeSceneH -> Ptr CDouble -> CInt -> CInt -> Ptr RenderOptsC -> Ptr CDouble -> Ptr Word32 -> Ptr () -> IO CInt
foreign import ccall safe   "glome_rayint_batch"   c_rayint_batch   :: Ptr GlomeSceneH -> Int64 -> Ptr CDouble -> Ptr CDouble -> CInt -> Ptr CudaHit -> IO CInt
foreign import ccall safe   "glome_shadow_batch"   c_shadow_batch   :: Ptr GlomeSceneH -> Int64 -> Ptr CDouble -> Ptr CDouble -> CInt -> Ptr Word8 -> IO CInt
foreign import ccall safe   "glome_trace_batch"    c_trace_batch    :: Ptr GlomeSceneH -> Int64 -> Ptr CDouble -> Ptr CDouble -> CInt -> CInt -> Ptr CDouble -> Ptr CDouble -> Ptr CudaHit -> IO CInt
foreign import ccall unsafe "glome_last_error"     c_last_error     :: IO CString
foreign import ccall unsafe "glome_render_opts_default" c_opts_default :: Ptr RenderOptsC -> IO ()
-- getTags' (Glome.hs:410-414): tags of the object under a pixel
foreign import ccall safe   "glome_get_tags"       c_get_tags       :: Ptr GlomeSceneH -> Ptr CDouble -> CInt -> CInt -> CInt -> CInt -> CInt -> Ptr Int32 -> CInt -> Ptr CInt -> Ptr CInt -> Ptr CudaHit -> IO CInt
-- scene set-up on the GPU: the trees of `bih` / `mesh` (Bih.hs:211-285, Mesh.hs:69-113), same arrays as the host builders
foreign import ccall safe   "glome_builder_set_build_device" c_builder_set_build_device :: Ptr GlomeBuilderH -> CInt -> IO CInt
-- several GPUs driven by this process: tile i on device (i mod n), gathered on the first device by peer copies
foreign import ccall safe   "glome_multi_create"   c_multi_create   :: Ptr GlomeFlatSceneC -> CInt -> Ptr CInt -> Ptr (Ptr GlomeMultiH) -> IO CInt
foreign import ccall safe   "glome_multi_destroy"  c_multi_destroy  :: Ptr GlomeMultiH -> IO CInt
foreign import ccall safe   "glome_multi_render"   c_multi_render   :: Ptr GlomeMultiH -> Ptr CDouble -> CInt -> CInt -> Ptr RenderOptsC -> Ptr CDouble -> Ptr Word32 -> Ptr () -> IO CInt
-- NFF / SPD scene text (Spd.hs)
foreign import ccall safe   "glome_sb_load_nff"    c_sb_load_nff    :: Ptr GlomeBuilderH -> CString -> Int64 -> Ptr CDouble -> Ptr CDouble -> Ptr Int64 -> IO CInt

foreign import ccall unsafe "glome_builder_create"  c_builder_create  :: Ptr (Ptr GlomeBuilderH) -> IO CInt
foreign import ccall unsafe "glome_builder_destroy" c_builder_destroy :: Ptr GlomeBuilderH -> IO CInt
foreign import ccall unsafe "glome_sb_flatten"      c_sb_flatten      :: Ptr GlomeBuilderH -> CInt -> Ptr GlomeFlatSceneC -> IO CInt
foreign import ccall unsafe "glome_sb_sphere"       c_sb_sphere       :: Ptr GlomeBuilderH -> Ptr CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_box"          c_sb_box          :: Ptr GlomeBuilderH -> Ptr CDouble -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_plane_offset" c_sb_plane_offset :: Ptr GlomeBuilderH -> Ptr CDouble -> CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_triangle"     c_sb_triangle     :: Ptr GlomeBuilderH -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "glome_sb_group"        c_sb_group        :: Ptr GlomeBuilderH -> CInt -> Ptr Int32 -> IO CInt
foreign import ccall unsafe "glome_sb_tex"          c_sb_tex          :: Ptr GlomeBuilderH -> CInt -> CInt -> IO CInt
foreign import ccall unsafe "glome_sb_tag"          c_sb_tag          :: Ptr GlomeBuilderH -> CInt -> CInt -> IO CInt
-- ... one import per glome_sb_* constructor; the remaining ones follow the same pattern
-- (cylinder_z, cone_z, disc, trianglenorm, difference, intersection, noshadow, onlyshadow,
--  bound_object, innerbound, mesh, materials, textures, lights).

-- | The additive patch to GlomeTrace (INTEGRATION.md): a new method that hands each Solid's
-- *already built* structure to the flattener.  `Bih` and `Mesh` pass their finished trees
-- (BihNode / BVH) rather than rebuilding them, so the device traverses exactly the tree the CPU
-- path would have traversed.  Closure textures cannot be introspected (Solid.hs:97), so the user
-- names them with the reified vocabulary of glome_cuda.h (GLOME_TEX_UNIFORM / _STRIPE_BLEND /
-- _PERLIN_BLEND) when calling `texCuda`.
newtype FlatBuilder = FlatBuilder (Ptr GlomeBuilderH)

class Flatten s where
  -- | Emit this solid into the builder, returning its item id.
  flatten :: FlatBuilder -> s -> IO Int

-- ---------------------------------------------------------------------------------------------
-- scene handle
-- ---------------------------------------------------------------------------------------------
-- | A scene resident on one GPU.  The C handle may be used by one host thread at a time
-- (glome_cuda.h), hence the MVar; the finalizer frees the device memory.
newtype CudaScene = CudaScene (MVar (ForeignPtr GlomeSceneH))

check :: CInt -> IO ()
check rc = when (rc < 0) $ do
  msg <- c_last_error >>= peekCString
  throwIO (ErrorCall ("GlomeTrace.CUDA: " ++ msg))   -- mirrors the reference's use of `error`

newCudaScene :: Flatten s => s -> Int -> IO CudaScene
newCudaScene sld device =
  alloca $ \pb -> do
    c_builder_create pb >>= check
    b <- peek pb
    root <- flatten (FlatBuilder b) sld
    sc <- allocaBytes 256 $ \fs -> do            -- sizeof(GlomeFlatScene) <= 256
      c_sb_flatten b (fromIntegral root) fs >>= check
      alloca $ \ph -> do
        c_scene_create fs (fromIntegral device) ph >>= check
        peek ph
    _ <- c_builder_destroy b                      -- glome_scene_create copied everything
    fp <- newForeignPtr p_scene_destroy sc
    CudaScene <$> newMVar fp

withCudaScene :: CudaScene -> (Ptr GlomeSceneH -> IO a) -> IO a
withCudaScene (CudaScene mv) act = withMVar mv $ \fp -> withForeignPtr fp act

-- ---------------------------------------------------------------------------------------------
-- renderTiles replacement (GlomeView/Glome.hs:379-386)
-- ---------------------------------------------------------------------------------------------
data RenderMode = OneRayPerPixel | AdaptiveAA deriving (Eq, Show)

data RenderOpts = RenderOpts
  { roMode       :: RenderMode
  , roBlocksize  :: Int        -- ^ 65 (Glome.hs:116)
  , roRecurs     :: Int        -- ^ maxdepth = 3 (Glome.hs:25)
  , roThresholds :: (Flt, Flt, Flt, Flt)  -- ^ 0.14 0.15 0.16 0.18 (Glome.hs:221-224)
  }

defaultRenderOpts :: RenderOpts
defaultRenderOpts = RenderOpts AdaptiveAA 65 3 (0.14, 0.15, 0.16, 0.18)

data RenderOptsC  -- struct GlomeRenderOpts, 64 bytes

pokeOpts :: Ptr RenderOptsC -> RenderOpts -> IO ()
pokeOpts p o = do
  c_opts_default p
  pokeByteOff p 0  (if roMode o == AdaptiveAA then 1 else 0 :: Int32)
  pokeByteOff p 4  (fromIntegral (roBlocksize o) :: Int32)
  pokeByteOff p 8  (fromIntegral (roRecurs o) :: Int32)
  let (a, b, c, d) = roThresholds o
  pokeByteOff p 16 a >> pokeByteOff p 24 b >> pokeByteOff p 32 c >> pokeByteOff p 40 d

-- | Drop-in for `renderTiles`: returns the frame as (r,g,b,a,depth) per pixel (the `Tile` payload,
-- Glome.hs:153-154) and as packed 0x00RRGGBB words (what `blitTile` writes to the SDL surface).
renderTilesCuda :: CudaScene -> Camera -> Int -> Int -> RenderOpts -> IO ([Flt], [Word32])
renderTilesCuda scn (Camera (Vec px py pz) (Vec fx fy fz) (Vec ux uy uz) (Vec rx ry rz)) w h o =
  withCudaScene scn $ \s ->
  withArray (map realToFrac [px,py,pz, fx,fy,fz, ux,uy,uz, rx,ry,rz]) $ \cam ->
  allocaBytes 64 $ \po ->
  allocaArray (5*w*h) $ \tc ->
  allocaArray (w*h) $ \rgb -> do
    pokeOpts po o
    c_render s cam (fromIntegral w) (fromIntegral h) po tc rgb nullPtr >>= check
    tcs <- map realToFrac <$> peekArray (5*w*h) tc
    px8 <- peekArray (w*h) rgb
    return (tcs, px8)

-- ---------------------------------------------------------------------------------------------
-- batch forms of rayint / shadow / trace (Solid.hs:146-162, Trace.hs:59)
-- ---------------------------------------------------------------------------------------------
-- | struct GlomeHit (144 bytes): the Rayint record as plain data.
data CudaHit = CudaHit { chDepth :: !Flt, chPos :: !Vec, chNorm :: !Vec, chHit :: !Bool
                       , chPrim :: !Int, chSub :: !Int, chTex :: [Int], chTag :: [Int] }

instance Storable CudaHit where
  sizeOf _ = 144
  alignment _ = 8
  peek p = do
    t <- peekByteOff p 0 :: IO CDouble
    [x,y,z,nx,ny,nz] <- mapM (\i -> realToFrac <$> (peekByteOff p (8+8*i) :: IO CDouble)) [0..5]
    hitf <- peekByteOff p 56 :: IO Int32
    prim <- peekByteOff p 60 :: IO Int32
    sub  <- peekByteOff p 64 :: IO Int32
    ntex <- peekByteOff p 68 :: IO Int32
    ntag <- peekByteOff p 72 :: IO Int32
    texs <- mapM (\i -> peekByteOff p (80+4*i) :: IO Int32) [0 .. fromIntegral ntex - 1]
    tags <- mapM (\i -> peekByteOff p (112+4*i) :: IO Int32) [0 .. fromIntegral ntag - 1]
    return (CudaHit (realToFrac t) (Vec x y z) (Vec nx ny nz) (hitf /= 0) (fromIntegral prim) (fromIntegral sub)
                    (map fromIntegral texs) (map fromIntegral tags))
  poke _ _ = error "CudaHit is read-only"

raysToC :: [Ray] -> [CDouble]
raysToC rs = concat [ map realToFrac [ox,oy,oz,dx,dy,dz] | Ray (Vec ox oy oz) (Vec dx dy dz) <- rs ]

-- | @rayint sld ray d [] []@ for a batch of rays.
rayintBatch :: CudaScene -> [Ray] -> Flt -> IO [CudaHit]
rayintBatch scn rays d = withCudaScene scn $ \s ->
  withArrayLen (raysToC rays) $ \_ pr -> with (realToFrac d) $ \pd ->
  allocaArray n $ \out -> do
    c_rayint_batch s (fromIntegral n) pr pd 0 out >>= check
    peekArray n out
  where n = length rays

-- | @shadow sld ray d@ for a batch of rays.
shadowBatch :: CudaScene -> [Ray] -> Flt -> IO [Bool]
shadowBatch scn rays d = withCudaScene scn $ \s ->
  withArrayLen (raysToC rays) $ \_ pr -> with (realToFrac d) $ \pd ->
  allocaArray n $ \out -> do
    c_shadow_batch s (fromIntegral n) pr pd 0 out >>= check
    map (/= 0) <$> peekArray n out
  where n = length rays

-- | @trace lights materialShader sld ray depth recurs@ for a batch of rays: (r,g,b,a) and ridepth.
traceBatch :: CudaScene -> [Ray] -> Flt -> Int -> IO [((Flt,Flt,Flt,Flt), Flt)]
traceBatch scn rays d recurs = withCudaScene scn $ \s ->
  withArrayLen (raysToC rays) $ \_ pr -> with (realToFrac d) $ \pd ->
  allocaArray (4*n) $ \rgba -> allocaArray n $ \dep -> do
    c_trace_batch s (fromIntegral n) pr pd 0 (fromIntegral recurs) rgba dep nullPtr >>= check
    cs <- map realToFrac <$> peekArray (4*n) rgba
    ds <- map realToFrac <$> peekArray n dep
    return (zip (quads cs) ds)
  where n = length rays
        quads (a:b:c:e:r) = (a,b,c,e) : quads r
        quads _ = []
