-- GlomeHeadless.hs -- headless timing harness for the REFERENCE (jimsnow/glome), to be dropped next
-- to GlomeView/Glome.hs and built with the reference's own flags:
--
--   ghc -O2 -threaded -rtsopts GlomeHeadless.hs && ./GlomeHeadless 720 480 +RTS -N<cores>
--
-- NOT RUN IN THIS REPOSITORY: the build image has no GHC (SURVEY.md F2), so BASELINE numbers quoted
-- here come from oracle/ (a C++ restatement).  This file exists so the true `+RTS -N` figure can be
-- produced wherever GHC is available.  It mirrors main/renderTiles (Glome.hs:440-469, 379-386) without
-- SDL: same 65x65 tiles, same parMap, same renderTileSubsample, and it forces every tile.
--
-- renderTile / renderTileSubsample / Tile are not exported by Glome.hs (it is a Main module); paste
-- this file's `main` into a copy of Glome.hs in place of the SDL `main`, keeping its imports.
import Control.DeepSeq (deepseq)
import Control.Monad.Par (runPar, parMap)
import Data.Time.Clock.POSIX (getPOSIXTime)
import System.Environment (getArgs)
import Graphics.UI.SDL (Rect(..))
import Data.Glome.Scene
import TestScene

blocksize :: Int
blocksize = 65

chunk :: Int -> Int -> [(Int, Int)]
chunk size bs = go 0
  where go pos | pos + bs >= size = [(pos, size - pos)]
               | otherwise        = (pos, bs) : go (pos + bs)

main :: IO ()
main = do
  [ws, hs] <- getArgs
  let (w, h) = (read ws, read hs) :: (Int, Int)
  scene@(geom, _, _, _) <- scn
  t0 <- getPOSIXTime
  print (primcount geom)                       -- forces BIH construction (Glome.hs:451)
  t1 <- getPOSIXTime
  let srect  = Rect 0 0 w h
      blocks = [Rect x y bw bh | (x, bw) <- chunk w blocksize, (y, bh) <- chunk h blocksize]
      tiles  = runPar $ parMap (\b -> renderTileSubsample srect b scene) blocks
  tiles `deepseq` return ()
  t2 <- getPOSIXTime
  putStrLn $ "setup_s " ++ show (t1 - t0) ++ " render_s " ++ show (t2 - t1)
          ++ " fps " ++ show (1 / realToFrac (t2 - t1) :: Double)
