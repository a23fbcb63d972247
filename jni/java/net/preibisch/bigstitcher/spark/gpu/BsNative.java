package net.preibisch.bigstitcher.spark.gpu;

import java.nio.ByteBuffer;

/**
 * Native methods of libbsgpu_jni.so (jni/bs_jni.cpp), one per entry point of include/bsgpu.h.
 * A context ({@code long ctx}) is bound to one device and one stream; create one per worker thread
 * (the Spark RDD of pairs / blocks collapses to a plain work queue over contexts).
 * Every failing call throws RuntimeException(bs_last_error).
 */
public final class BsNative
{
	static { System.loadLibrary( "bsgpu_jni" ); }

	private BsNative() {}

	public static final int U16 = 0, F32 = 1, U8 = 2;
	/** FusionType ordinals (mvrecon FusionGUI.FusionType) */
	public static final int AVG = 0, AVG_BLEND = 1, AVG_CONTENT = 2, AVG_BLEND_CONTENT = 3, MAX_INTENSITY = 4,
			LOWEST_VIEWID_WINS = 5, HIGHEST_VIEWID_WINS = 6, CLOSEST_PIXEL_WINS = 7;
	/** layout of the double[] a pcm* call returns per pair */
	public static final int R_FOUND = 0, R_SHIFT_INT = 1, R_SHIFT_SUB = 4, R_R = 7, R_N_OVERLAP = 8, R_PEAK = 9,
			R_PCM_VALUE = 12, R_PAD = 13, R_N_CANDIDATES = 16, R_PEARSON_PX = 17, R_STRIDE = 20;

	public static native int version();
	public static native long init( int device );
	public static native void destroy( long ctx );
	public static native String lastError( long ctx );
	public static native void synchronize( long ctx );
	public static native long launchCount( long ctx );
	public static native void profileEnable( long ctx, boolean on );
	public static native void profileReset( long ctx );
	public static native double[] profileGet( long ctx, String tag );
	public static native ByteBuffer hostAlloc( long ctx, long bytes );
	public static native void hostFree( long ctx, ByteBuffer buffer );

	/** data: short[] / float[] / byte[] or a direct ByteBuffer; dims {x,y,z} */
	public static native long volumeUpload( long ctx, Object data, long[] dims, int dtype );
	public static native long volumeUploadAsync( long ctx, ByteBuffer pinned, long[] dims, int dtype );
	public static native long volumeWrap( long ctx, long devicePointer, long[] dims, int dtype );
	public static native void volumeFree( long ctx, long handle );
	/** {dimX, dimY, dimZ, dtype} */
	public static native long[] volumeInfo( long ctx, long handle );
	public static native long volumeDevptr( long ctx, long handle );
	public static native void volumeDownload( long ctx, long handle, Object dest, long capacityBytes );
	public static native long contentWeights( long ctx, long handle, double sigma1, double sigma2 );
	public static native long downsample( long ctx, long handle, int[] factors );

	public static native int goodFftSize( int n, boolean even );
	/** iparams {peaksToCheck, doSubpixel, extX, extY, extZ} */
	public static native double[] pcmPair( long ctx, Object img1, Object img2, long[] dims, int dtype, int[] iparams, double minOverlap );
	public static native double[] pcmBatch( long ctx, Object[] imgs1, Object[] imgs2, long[] dims, int dtype, int[] iparams, double minOverlap );
	/** jobs: n x {vol1, vol2, min1[3], min2[3], dims[3]} */
	public static native double[] pcmVolumesBatch( long ctx, long[] jobs, int[] iparams, double minOverlap );
	public static native long[] pcmDebugPcm( long ctx, Object img1, Object img2, long[] dims, int dtype, int[] extension, Object outPcm );

	/** models n*12, handles n*{volume, content}, blend n*{border[3], range[3]}, windows n*{fullDims[3], windowMin[3]} or null;
	 *  iparams {fusionType, interpolation, outDtype, blendLutN, outBigEndian (1: N5 block payload byte order)}; dparams {minIntensity, maxIntensity} */
	public static native void fuseBlock( long ctx, int nViews, double[] models, long[] handles, float[] blend, long[] windows,
			long[] blockMin, long[] blockSize, int[] iparams, double[] dparams, Object dest );
	public static native void fuseBlocks( long ctx, int nViews, double[] models, long[] handles, float[] blend, long[] windows,
			long[] blockMins, long[] blockSizes, int[] iparams, double[] dparams, Object[] dests );
	/** --masks: 255 / 65535 / 1.0f where any view's (grown) pixel grid covers the voxel (GenerateComputeBlockMasks) */
	public static native void maskBlocks( long ctx, int nViews, double[] models, long[] handles, long[] windows,
			long[] blockMins, long[] blockSizes, double[] maskOffset, int outDtype, int outBigEndian, Object[] dests );
	public static native long fuseBlockToVolume( long ctx, int nViews, double[] models, long[] handles, float[] blend, long[] windows,
			long[] blockMin, long[] blockSize, int[] iparams, double[] dparams );
	public static native void fuseAccumulate( long ctx, int nViews, double[] models, long[] handles, float[] blend,
			long[] blockMin, long[] blockSize, int[] iparams, double[] dparams, long sumWiDev, long sumWDev );
	/** view-sharded exchange: rank 0 draws the id, every rank joins, fuseAllreduce sums both partial buffers over NVLink */
	public static native byte[] commUniqueId();
	public static native void commInit( long ctx, int nRanks, int rank, byte[] id );
	public static native void commDestroy( long ctx );
	public static native void fuseAllreduce( long ctx, long sumWiDev, long sumWDev, long n );
	public static native void fuseFinish( long ctx, long sumWiDev, long sumWDev, long n, int[] iparams, double[] dparams, Object dest );

	/** n x {locX, locY, locZ, value, voxelX, voxelY, voxelZ, isMax}; dparams {sigma, threshold, minI, maxI}; iparams {findMax, findMin, localization} */
	public static native double[] dogDetect( long ctx, long handle, long[] intervalMin, long[] intervalSize, double[] dparams, int[] iparams );
}
