{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE EmptyDataDecls #-}
-- | GlomeTrace.CUDA: the reference-side binding to libglomecuda.so (include/glome_cuda.h).
--
-- NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no GHC (SURVEY.md F2).  This module, its bottom layer
-- "Data.Glome.CUDA.Flat" and the patch haskell/glometrace-cuda.patch (the `flatten` method of class `Solid` with one
-- instance per solid: Sphere, Triangle, TriangleNorm, Box, Plane, Disc, Cylinder, Cone, [s], Void, Instance, Bih, Mesh,
-- Difference, Intersection, Tex, TexD, Tag, NoShadow, OnlyShadow, Bound, InnerBound) are what a maintainer adds to the
-- GlomeTrace package (INTEGRATION.md).  Only `foreign import ccall` over the C shim is used, as BASELINE.json's
-- north_star asks.
--
-- Scene construction stays in Haskell: build the scene with the usual GlomeTrace constructors (sphere, box, cone,
-- difference, bih, mesh, tex / texD, transform ...), call 'newCudaScene' once, then render frames / trace ray batches
-- on the B200.  `Bih` and `Mesh` hand over the trees their constructors built; nothing is rebuilt on the C side.
module Data.Glome.CUDA
  ( CudaScene, CudaReify(..), defaultReify, TexDesc(..)
  , withCudaScene, newCudaScene
  , renderTilesCuda, RenderMode(..), RenderOpts(..), defaultRenderOpts
  , rayintBatch, shadowBatch, traceBatch, getTagsCuda
  , CudaHit(..)
  ) where

import Control.Concurrent.MVar
import Control.Exception (throwIO, ErrorCall(..))
import Control.Monad (when, forM, forM_)
import Data.IORef
import Data.Int
import qualified Data.Map as M
import Data.Word
import Foreign
import Foreign.C.String
import Foreign.C.Types
import System.Mem.StableName

import Data.Glome.Vec
import Data.Glome.Clr
import Data.Glome.Solid
import Data.Glome.Scene (Camera(..))
import Data.Glome.Shader (Material(..), Light(..))
import Data.Glome.CUDA.Flat

-- ---------------------------------------------------------------------------------------------
-- C side (include/glome_cuda.h); the scene-construction imports live in Data.Glome.CUDA.Flat
-- ---------------------------------------------------------------------------------------------
data GlomeSceneH     -- opaque GlomeScene
data GlomeMultiH

-- calls block on the GPU: import them `safe` so other Haskell threads keep running
foreign import ccall safe   "glome_scene_create"   c_scene_create   :: Ptr () -> CInt -> Ptr (Ptr GlomeSceneH) -> IO CInt
foreign import ccall safe   "glome_scene_destroy"  c_scene_destroy  :: Ptr GlomeSceneH -> IO CInt
foreign import ccall safe   "&glome_scene_destroy" p_scene_destroy  :: FunPtr (Ptr GlomeSceneH -> IO ())
foreign import ccall safe   "glome_render"         c_render         :: Ptr GlomeSceneH -> Ptr CDouble -> CInt -> CInt -> Ptr RenderOptsC -> Ptr CDouble -> Ptr Word32 -> Ptr () -> IO CInt
foreign import ccall safe   "glome_rayint_batch"   c_rayint_batch   :: Ptr GlomeSceneH -> Int64 -> Ptr CDouble -> Ptr CDouble -> CInt -> Ptr CudaHit -> IO CInt
foreign import ccall safe   "glome_shadow_batch"   c_shadow_batch   :: Ptr GlomeSceneH -> Int64 -> Ptr CDouble -> Ptr CDouble -> CInt -> Ptr Word8 -> IO CInt
foreign import ccall safe   "glome_trace_batch"    c_trace_batch    :: Ptr GlomeSceneH -> Int64 -> Ptr CDouble -> Ptr CDouble -> CInt -> CInt -> Ptr CDouble -> Ptr CDouble -> Ptr CudaHit -> IO CInt
foreign import ccall unsafe "glome_render_opts_default" c_opts_default :: Ptr RenderOptsC -> IO ()
-- getTags' (Glome.hs:410-414): tags of the object under a pixel
foreign import ccall safe   "glome_get_tags"       c_get_tags       :: Ptr GlomeSceneH -> Ptr CDouble -> CInt -> CInt -> CInt -> CInt -> CInt -> Ptr Int32 -> CInt -> Ptr CInt -> Ptr CInt -> Ptr CudaHit -> IO CInt
-- several GPUs driven by this process: tile i on device (i mod n), gathered on the first device by peer copies
foreign import ccall safe   "glome_multi_create"   c_multi_create   :: Ptr () -> CInt -> Ptr CInt -> Ptr (Ptr GlomeMultiH) -> IO CInt
foreign import ccall safe   "glome_multi_destroy"  c_multi_destroy  :: Ptr GlomeMultiH -> IO CInt
foreign import ccall safe   "glome_multi_render"   c_multi_render   :: Ptr GlomeMultiH -> Ptr CDouble -> CInt -> CInt -> Ptr RenderOptsC -> Ptr CDouble -> Ptr Word32 -> Ptr () -> IO CInt
-- NFF / SPD scene text (Spd.hs)
foreign import ccall safe   "glome_sb_load_nff"    c_sb_load_nff    :: Ptr GlomeBuilderH -> CString -> Int64 -> Ptr CDouble -> Ptr CDouble -> Ptr Int64 -> IO CInt

-- ---------------------------------------------------------------------------------------------
-- what Haskell cannot look into: closures
-- ---------------------------------------------------------------------------------------------
-- | `Texture t m`, a light's falloff and Warp's ray transform are functions (Solid.hs:97, Shader.hs:16, 47-50).  The
-- device evaluates the reified vocabulary of glome_cuda.h, so the caller says which closure is which:
--
--  * textures attached with 'Data.Glome.Shader.texD' carry their own description and need nothing here;
--  * a plain @tex s f@ is resolved through 'reifyTexture' (typically a lookup of @f@'s StableName in a table the
--    scene file fills: @[(t_stripe, TexStripeBlend m_shiny_white m_dull_gray (Vec 4 8 5)), ...]@, see 'textureTable');
--  * the n-th Warp material met during flattening gets its ray transform from 'reifyWarp' (TestScene.hs:169-173 is
--    @xfm_ray X (Ray pos (vnorm dir))@: return that X);
--  * lights are assumed to have the inverse-square falloff of `light` (Shader.hs:22).
data CudaReify t = CudaReify
  { reifyTexture :: Texture t (Material t) -> IO (Maybe (TexDesc (Material t)))
  , reifyWarp    :: Int -> Maybe Xfm
  }

defaultReify :: CudaReify t
defaultReify = CudaReify (\_ -> return Nothing) (const Nothing)

-- | A 'reifyTexture' from a table of (closure, description) pairs, compared by StableName: the table must hold the
-- very closures the scene was built with (top-level definitions such as t_stripe are).
textureTable :: [(Texture t (Material t), TexDesc (Material t))] -> IO (Texture t (Material t) -> IO (Maybe (TexDesc (Material t))))
textureTable pairs = do
  keyed <- forM pairs $ \(f, d) -> do { n <- makeStableName f; return (n, d) }
  return $ \f -> do n <- makeStableName f
                    return (lookup n keyed)

-- ---------------------------------------------------------------------------------------------
-- flattening a scene: solids (class method `flatten`), materials, textures, tags, lights
-- ---------------------------------------------------------------------------------------------
data FlatState t = FlatState
  { fsTags    :: IORef (M.Map t Int)                                 -- tag value -> dense id
  , fsWarps   :: IORef Int                                           -- Warp materials met so far
  , fsPending :: IORef [(Int, SolidItem t (Material t))]             -- Warp material id -> its scene, flattened last
  }

-- | Materials (Shader.hs:38-50) -> glome_sb_mat_*.  A Warp's scene may be the scene that contains it
-- (TestScene.hs:179 re-casts into geom'' itself), so it is not flattened here: it is queued and resolved after the root
-- (glome_sb_mat_warp_set_scene).
flattenMaterial :: Ord t => CudaReify t -> FlatState t -> FlatBuilder (Texture t (Material t)) t (Material t) -> Material t -> IO Int
flattenMaterial rf st fb mat = case mat of
  Surface (Color r g b) alpha amb kd ks shine _ ->
    withArray (map realToFrac [r, g, b]) $ \p ->
      c_sb_mat_surface h p (realToFrac alpha) (realToFrac amb) (realToFrac kd) (realToFrac ks) (realToFrac shine) >>= chk "Surface"
  Reflect a     -> c_sb_mat_reflect h (realToFrac a) >>= chk "Reflect"
  Refract a b c -> c_sb_mat_refract h (realToFrac a) (realToFrac b) (realToFrac c) >>= chk "Refract"
  Blend a b w   -> do ia <- flattenMaterial rf st fb a
                      ib <- flattenMaterial rf st fb b
                      c_sb_mat_blend h (fromIntegral ia) (fromIntegral ib) (realToFrac w) >>= chk "Blend"
  AdditiveLayers ms -> do ids <- mapM (flattenMaterial rf st fb) ms
                          withArrayLen (map fromIntegral ids) $ \n p -> c_sb_mat_additive h (fromIntegral n) p >>= chk "AdditiveLayers"
  Warp frame scene lights _ -> do
    k <- atomicModifyIORef (fsWarps st) (\n -> (n + 1, n))
    x <- maybe (throwIO (ErrorCall ("GlomeTrace.CUDA: Warp #" ++ show k ++ ": reifyWarp gave no ray transform"))) return (reifyWarp rf k)
    iframe <- flatten fb frame
    ls <- flattenLights fb lights
    placeholder <- emitVoid fb
    m <- withArray (xfm24 x) $ \p -> c_sb_mat_warp h (fromIntegral iframe) (fromIntegral placeholder) (fromIntegral ls) p >>= chk "Warp"
    modifyIORef (fsPending st) ((m, scene) :)
    return m
 where h = fbHandle fb

flattenTexDesc :: Ord t => CudaReify t -> FlatState t -> FlatBuilder (Texture t (Material t)) t (Material t) -> TexDesc (Material t) -> IO Int
flattenTexDesc rf st fb d = case d of
  TexUniform m -> do i <- flattenMaterial rf st fb m
                     c_sb_tex_uniform h (fromIntegral i) >>= chk "t_uniform"
  TexStripeBlend a b axis -> do ia <- flattenMaterial rf st fb a
                                ib <- flattenMaterial rf st fb b
                                withArray (vec3 axis) $ \p -> c_sb_tex_stripe_blend h (fromIntegral ia) (fromIntegral ib) p >>= chk "stripe"
  TexPerlinBlend a b s -> do ia <- flattenMaterial rf st fb a
                             ib <- flattenMaterial rf st fb b
                             c_sb_tex_perlin_blend h (fromIntegral ia) (fromIntegral ib) (realToFrac s) >>= chk "perlin"
 where h = fbHandle fb

-- | `[Light]` (Shader.hs:12-23) -> one light set.  Inverse-square falloff, infinite radius, shadows on (`light`).
flattenLights :: FlatBuilder a b c -> [Light] -> IO Int
flattenLights fb ls = do
  ids <- forM ls $ \l -> case litcol l of
    Color r g b -> withArray (vec3 (litpos l)) $ \pp -> withArray (map realToFrac [r, g, b]) $ \pc ->
                     c_sb_light (fbHandle fb) pp pc >>= chk "light"
  withArrayLen (map fromIntegral ids) $ \n p -> c_sb_lightset (fbHandle fb) (fromIntegral n) p >>= chk "lights"

mkBuilder :: Ord t => CudaReify t -> Ptr GlomeBuilderH -> IO (FlatState t, FlatBuilder (Texture t (Material t)) t (Material t))
mkBuilder rf h = do
  st <- FlatState <$> newIORef M.empty <*> newIORef 0 <*> newIORef []
  let fb = FlatBuilder
        { fbHandle  = h
        , fbTexture = \f -> do md <- reifyTexture rf f
                               case md of
                                 Just d  -> flattenTexDesc rf st fb d
                                 Nothing -> throwIO (ErrorCall "GlomeTrace.CUDA: a texture closure with no description (use texD, or reifyTexture)")
        , fbTexDesc = flattenTexDesc rf st fb
        , fbTag     = \t -> atomicModifyIORef (fsTags st) (\m -> case M.lookup t m of
                                                                    Just i  -> (m, i)
                                                                    Nothing -> let i = M.size m in (M.insert t i m, i))
        }
  return (st, fb)

-- ---------------------------------------------------------------------------------------------
-- scene handle
-- ---------------------------------------------------------------------------------------------
-- | A scene resident on one GPU.  The C handle may be used by one host thread at a time
-- (glome_cuda.h), hence the MVar; the finalizer frees the device memory.
data CudaScene t = CudaScene (MVar (ForeignPtr GlomeSceneH)) (M.Map Int t)   -- dense device tag id -> the scene's tag

check :: CInt -> IO ()
check rc = chk "call" rc >> return ()             -- mirrors the reference's use of `error`

-- | Flatten @(geometry, lights)@ -- the first two components of GlomeView's `Scene` -- and upload it to GPU @device@.
-- The scene's own light list must be emitted first: light set 0 is the one `trace` starts with.
newCudaScene :: (Ord t, Solid s t (Material t)) => CudaReify t -> s -> [Light] -> Int -> IO (CudaScene t)
newCudaScene rf sld lights device =
  alloca $ \pb -> do
    _ <- c_builder_create pb >>= chk "builder"
    b <- peek pb
    (st, fb) <- mkBuilder rf b
    _ <- flattenLights fb lights
    rootName <- sld `seq` makeStableName sld      -- pass the scene as the value the Warp holds (geom'' :: SolidItem)
    root <- flatten fb sld
    -- Warp scenes, last: the scene a portal looks into is usually the root itself
    let resolve = do
          pend <- atomicModifyIORef (fsPending st) (\p -> ([], p))
          forM_ pend $ \(m, scn) -> do
            n <- scn `seq` makeStableName scn
            i <- if eqStableName n rootName then return root else flatten fb scn
            c_sb_mat_warp_set_scene b (fromIntegral m) (fromIntegral i) >>= chk "Warp scene"
          when (not (null pend)) resolve
    resolve
    sc <- allocaBytes 256 $ \fs -> do            -- sizeof(GlomeFlatScene) <= 256 (asserted in glome_cuda.h)
      _ <- c_sb_flatten b (fromIntegral root) fs >>= chk "flatten"
      alloca $ \ph -> do
        _ <- c_scene_create fs (fromIntegral device) ph >>= chk "scene_create"
        peek ph
    _ <- c_builder_destroy b                      -- glome_scene_create copied everything
    tagmap <- readIORef (fsTags st)
    fp <- newForeignPtr p_scene_destroy sc
    mv <- newMVar fp
    return (CudaScene mv (M.fromList [ (i, t) | (t, i) <- M.toList tagmap ]))

withCudaScene :: CudaScene t -> (Ptr GlomeSceneH -> IO a) -> IO a
withCudaScene (CudaScene mv _) act = withMVar mv $ \fp -> withForeignPtr fp act

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

data RenderOptsC  -- struct GlomeRenderOpts, 64 bytes; the offsets below are GLOME_LAYOUT_ASSERTed in glome_cuda.h

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
renderTilesCuda :: CudaScene t -> Camera -> Int -> Int -> RenderOpts -> IO ([Flt], [Word32])
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
-- | struct GlomeHit (144 bytes): the Rayint record as plain data.  Offsets: GLOME_LAYOUT_ASSERTs of glome_cuda.h.
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
rayintBatch :: CudaScene t -> [Ray] -> Flt -> IO [CudaHit]
rayintBatch scn rays d = withCudaScene scn $ \s ->
  withArrayLen (raysToC rays) $ \_ pr -> with (realToFrac d) $ \pd ->
  allocaArray n $ \out -> do
    c_rayint_batch s (fromIntegral n) pr pd 0 out >>= check
    peekArray n out
  where n = length rays

-- | @shadow sld ray d@ for a batch of rays.
shadowBatch :: CudaScene t -> [Ray] -> Flt -> IO [Bool]
shadowBatch scn rays d = withCudaScene scn $ \s ->
  withArrayLen (raysToC rays) $ \_ pr -> with (realToFrac d) $ \pd ->
  allocaArray n $ \out -> do
    c_shadow_batch s (fromIntegral n) pr pd 0 out >>= check
    map (/= 0) <$> peekArray n out
  where n = length rays

-- | @trace lights materialShader sld ray depth recurs@ for a batch of rays: (r,g,b,a) and ridepth.
traceBatch :: CudaScene t -> [Ray] -> Flt -> Int -> IO [((Flt,Flt,Flt,Flt), Flt)]
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

-- | getTags' (Glome.hs:410-414): the tag list `ts ++ tags` (Trace.hs:82) of the trace through pixel (x, y), mapped back
-- from the device's dense ids to the scene's own tag values.
getTagsCuda :: CudaScene t -> Camera -> Int -> Int -> Int -> Int -> Int -> IO [t]
getTagsCuda scn@(CudaScene _ tagmap) (Camera (Vec px py pz) (Vec fx fy fz) (Vec ux uy uz) (Vec rx ry rz)) w h x y recurs =
  withCudaScene scn $ \s ->
  withArray (map realToFrac [px,py,pz, fx,fy,fz, ux,uy,uz, rx,ry,rz]) $ \cam ->
  allocaArray cap $ \out -> alloca $ \pn -> alloca $ \povf -> do
    c_get_tags s cam (fromIntegral w) (fromIntegral h) (fromIntegral x) (fromIntegral y) (fromIntegral recurs)
               out (fromIntegral cap) pn povf nullPtr >>= check
    n <- peek pn
    ids <- peekArray (fromIntegral n) out
    return [ t | i <- ids, Just t <- [M.lookup (fromIntegral i) tagmap] ]
  where cap = 512
