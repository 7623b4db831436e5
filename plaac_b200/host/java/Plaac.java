// Plaac.java -- Java host with plaac.jar's command line, parameter block and output tables; all scoring through
// libplaac_cuda.so (include/plaac_cuda.h) over the Foreign Function & Memory API (JDK >= 22, no JNI glue).
//
//     javac Plaac.java && LD_LIBRARY_PATH=<dir of libplaac_cuda.so> java --enable-native-access=ALL-UNNAMED Plaac -i in.fa
//
// SOURCE ONLY: the build image of this repository has no JDK, so this file has never been compiled there; the tested
// host with the same behaviour is plaac_b200/host/plaac_cli.cpp, of which this is the Java counterpart (same reader
// semantics, same parameter chain, same columns).  What it replaces in the reference: plaac.java main (:302-530,
// flags, background/foreground frequencies, llr and HMM tables -- computed HERE with java.lang.Math so the tables carry
// the JVM's bits), fastareader (:4302-4375), and the two driver loops scoreallfastas (:653-950) / plotsomefastas
// (:587-649) whose per-protein bodies become one plaac_score() call per batch.  There is no CPU scoring path.
import java.io.BufferedReader;
import java.io.FileReader;
import java.io.IOException;
import java.io.PrintStream;
import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

public final class Plaac {
    static final String AA = "XACDEFGHIKLMNPQRSTVWY*";   // residue alphabet, code = index (plaac.java:26)
    static final int NAA = 22, LUT = 4001;

    // ------------------------------------------------------------------------------------------ the C ABI
    static final class Cuda implements AutoCloseable {
        private static final Linker LINKER = Linker.nativeLinker();
        private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.mapLibraryName("plaac_cuda"), Arena.global());

        private static MethodHandle fn(String name, FunctionDescriptor d) {
            return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
        }

        private static final MethodHandle DEVICE_COUNT = fn("plaac_device_count", FunctionDescriptor.of(JAVA_INT));
        private static final MethodHandle CREATE = fn("plaac_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
        private static final MethodHandle DESTROY = fn("plaac_destroy", FunctionDescriptor.ofVoid(ADDRESS));
        private static final MethodHandle LAST_ERROR = fn("plaac_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
        // int plaac_score(ctx, codes, offsets, nprot, summaries, per_res)
        private static final MethodHandle SCORE =
            fn("plaac_score", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));

        // int plaac_score_multi(ctxs, nctx, codes, offsets, nprot, summaries, per_res): one ctx + host thread per GPU
        private static final MethodHandle SCORE_MULTI = fn("plaac_score_multi",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
        // transport-lean calls: radix-22 words (7 residues per int) + int lengths in, optional ranked hits out
        private static final MethodHandle PACKED_WORDS = fn("plaac_packed_words", FunctionDescriptor.of(JAVA_LONG, JAVA_LONG));
        private static final MethodHandle PACK_HOST =
            fn("plaac_pack_host", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT));
        private static final MethodHandle SCORE_PACKED = fn("plaac_score_packed",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
        private static final MethodHandle SCORE_MULTI_PACKED = fn("plaac_score_multi_packed",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
        // int plaac_rank(ctx, summaries, nprot, flags, order, n_core): the web front end's order (server.rb:222-229)
        private static final MethodHandle RANK =
            fn("plaac_rank", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS));
        // int plaac_score_fasta(ctx, text, nbytes, max_rec, summaries, codes, offsets, name_pos, name_len, flags, index, bg_counts)
        private static final MethodHandle SCORE_FASTA = fn("plaac_score_fasta", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS,
            JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));

        // int plaac_host_alloc(void **out, size_t bytes, int flags); int plaac_host_free(void *p)
        private static final MethodHandle HOST_ALLOC =
            fn("plaac_host_alloc", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT));
        private static final MethodHandle HOST_FREE = fn("plaac_host_free", FunctionDescriptor.of(JAVA_INT, ADDRESS));

        static final long SUMMARY_BYTES = 160;   // plaac_summary: 14 x int32, 13 x double
        // plaac_params: 8 x int32, then lt[2][2] li[2] lf[2] le[2][22] le0 llr papa_lod hydro2 charge (22 each) fi_cc[3]
        // big_neg ln2 loglut[4001]
        static final long PARAMS_BYTES = 8 * 4 + 8L * (4 + 2 + 2 + 2 * NAA + 5 * NAA + 3 + 1 + 1 + LUT);

        private final Arena arena = Arena.ofConfined();
        private final MemorySegment ctx;

        static int deviceCount() throws Throwable {
            return (int) DEVICE_COUNT.invokeExact();
        }

        Cuda(int device, Params p) throws Throwable {
            MemorySegment s = arena.allocate(PARAMS_BYTES, 8);
            long o = 0;
            for (int v : new int[] {p.coreLength, p.ww1, p.ww2, p.ww3, p.adjustProlines ? 1 : 0, 80, 0, 0}) {
                s.set(JAVA_INT, o, v);
                o += 4;
            }
            o = put(s, o, p.lt[0]);
            o = put(s, o, p.lt[1]);
            o = put(s, o, p.li);
            o = put(s, o, p.lf);
            o = put(s, o, p.le[0]);
            o = put(s, o, p.le[1]);
            o = put(s, o, p.le[0]);          // hmm0 emits the background in both states (prionhmm0)
            o = put(s, o, p.llr);
            o = put(s, o, p.papaLod);
            o = put(s, o, p.hydro2);
            o = put(s, o, p.charge);
            o = put(s, o, new double[] {2.785, -1, -1.151});
            o = put(s, o, new double[] {-1000000.0, Math.log(2.0)});
            o = put(s, o, p.logLut);
            MemorySegment out = arena.allocate(ADDRESS);
            int rc = (int) CREATE.invokeExact(out, device, s);
            if (rc != 0) throw new IllegalStateException("plaac_create failed (" + rc + "): " + lastError(MemorySegment.NULL));
            ctx = out.get(ADDRESS, 0);
        }

        private static long put(MemorySegment s, long o, double[] v) {
            for (double x : v) {
                s.set(JAVA_DOUBLE, o, x);
                o += 8;
            }
            return o;
        }

        static String lastError(MemorySegment c) throws Throwable {
            MemorySegment m = (MemorySegment) LAST_ERROR.invokeExact(c);
            return m.reinterpret(512).getString(0);
        }

        /** Page-locked native memory (plaac_host_alloc), released with the arena: plaac_score overlaps the copies of a
         *  batch with its compute only when the batch lies in pinned memory (pageable memory is staged by the driver). */
        static MemorySegment pinned(long bytes, Arena a) throws Throwable {
            long n = Math.max(1L, bytes);
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) HOST_ALLOC.invokeExact(out, n, 0);
            if (rc != 0) throw new IllegalStateException("plaac_host_alloc failed (" + rc + "): " + lastError(MemorySegment.NULL));
            return out.get(ADDRESS, 0).reinterpret(n, a, seg -> {
                try {
                    int ignored = (int) HOST_FREE.invokeExact(seg);
                } catch (Throwable t) {
                    throw new IllegalStateException(t);
                }
            });
        }

        /** Summary records of one batch (record i at i * 160). */
        MemorySegment score(Batch b, Arena a) throws Throwable {
            int nprot = b.names.size();
            MemorySegment codes = pinned(b.ncodes, a);
            MemorySegment.copy(b.codes, 0, codes, JAVA_BYTE, 0, b.ncodes);
            MemorySegment offs = pinned(8L * (nprot + 1), a);
            for (int i = 0; i <= nprot; i++) offs.setAtIndex(JAVA_LONG, i, b.offsets.get(i));
            MemorySegment sum = pinned(SUMMARY_BYTES * nprot, a);
            int rc = (int) SCORE.invokeExact(ctx, codes, offs, (long) nprot, sum, MemorySegment.NULL);
            if (rc != 0) throw new IllegalStateException("plaac_score failed (" + rc + "): " + lastError(ctx));
            return sum;
        }

        /** Summary records of one batch scored on several GPUs at once (--gpus N): the batch goes over the host link as
         *  radix-22 words (4/7 byte per residue) and int lengths, each ctx scores a residue-balanced contiguous shard
         *  and writes its records at the input position (plaac_score_multi_packed; no collective, no gather copy). */
        static MemorySegment scoreMulti(List<Cuda> ctxs, Batch b, Arena a) throws Throwable {
            int nprot = b.names.size();
            MemorySegment codes = a.allocate(Math.max(1, b.ncodes));
            MemorySegment.copy(b.codes, 0, codes, JAVA_BYTE, 0, b.ncodes);
            long nwords = (long) PACKED_WORDS.invokeExact((long) b.ncodes);
            MemorySegment words = pinned(4L * Math.max(1, nwords), a);
            int rc = (int) PACK_HOST.invokeExact(codes, (long) b.ncodes, words, 0);
            if (rc != 0) throw new IllegalStateException("plaac_pack_host failed (" + rc + ")");
            MemorySegment lens = pinned(4L * nprot, a);
            for (int i = 0; i < nprot; i++) lens.setAtIndex(JAVA_INT, i, (int) (b.offsets.get(i + 1) - b.offsets.get(i)));
            MemorySegment sum = pinned(SUMMARY_BYTES * nprot, a);
            MemorySegment handles = a.allocate(ADDRESS.byteSize() * ctxs.size(), 8);
            for (int k = 0; k < ctxs.size(); k++) handles.setAtIndex(ADDRESS, k, ctxs.get(k).ctx);
            rc = (int) SCORE_MULTI_PACKED.invokeExact(handles, ctxs.size(), words, lens, (long) nprot, (long) b.ncodes, sum,
                MemorySegment.NULL, MemorySegment.NULL);
            if (rc != 0) throw new IllegalStateException("plaac_score_multi_packed failed (" + rc + "): " + lastError(ctxs.get(0).ctx));
            return sum;
        }

        /** Row order of the reference's web front end (COREscore desc, LLR desc, rows without a CORE last): --rank. */
        int[] rank(MemorySegment summaries, int nprot, Arena a) throws Throwable {
            MemorySegment order = a.allocate(4L * Math.max(1, nprot), 4), ncore = a.allocate(8, 8);
            int rc = (int) RANK.invokeExact(ctx, summaries, (long) nprot, 0, order, ncore);
            if (rc != 0) throw new IllegalStateException("plaac_rank failed (" + rc + "): " + lastError(ctx));
            int[] out = new int[nprot];
            for (int i = 0; i < nprot; i++) out[i] = order.getAtIndex(JAVA_INT, i);
            return out;
        }

        /** Per-residue arrays of one batch: 2 byte arrays (vit, map) and 10 double arrays, each ncodes long. */
        MemorySegment[] scoreResidues(Batch b, Arena a) throws Throwable {
            int nprot = b.names.size();
            long n = Math.max(1, b.ncodes);
            MemorySegment codes = a.allocate(n);
            MemorySegment.copy(b.codes, 0, codes, JAVA_BYTE, 0, b.ncodes);
            MemorySegment offs = a.allocate(8L * (nprot + 1), 8);
            for (int i = 0; i <= nprot; i++) offs.setAtIndex(JAVA_LONG, i, b.offsets.get(i));
            MemorySegment[] arr = new MemorySegment[12];
            MemorySegment res = a.allocate(12L * ADDRESS.byteSize(), 8);   // plaac_residue_out: 12 pointers
            for (int k = 0; k < 12; k++) {
                arr[k] = k < 2 ? a.allocate(n) : a.allocate(8 * n, 8);
                res.setAtIndex(ADDRESS, k, arr[k]);
            }
            int rc = (int) SCORE.invokeExact(ctx, codes, offs, (long) nprot, MemorySegment.NULL, res);
            if (rc != 0) throw new IllegalStateException("plaac_score failed (" + rc + "): " + lastError(ctx));
            return arr;
        }

        @Override
        public void close() {
            try {
                DESTROY.invokeExact(ctx);
            } catch (Throwable t) {
                // nothing to do
            }
            arena.close();
        }
    }

    // ------------------------------------------------------------------------------------------ parameters
    /** Everything main derives between the flags and the scoring loops (plaac.java:310-518), with java.lang.Math. */
    static final class Params {
        int coreLength = 60, ww1 = 41, ww2 = 41, ww3 = 41;
        boolean adjustProlines = true;
        double alpha = 1.0;
        double[] fg, bgScer, bgInput, bg, llr = new double[NAA], papaLod = new double[NAA], hydro2 = new double[NAA];
        double[] charge = {0, 0, 0, 1, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0};
        double[][] lt = new double[2][2], le = new double[2][NAA];
        double[] li = new double[2], lf = new double[2], logLut = new double[LUT];

        static final double[] HYDRO = {0.0, 1.8, 2.5, -3.5, -3.5, 2.8, -0.4, -3.2, 4.5, -3.9, 3.8, 1.9, -3.5, -1.6, -3.5, -4.5,
            -0.8, -0.7, 4.2, -0.9, -1.3, 0.0};
        static final double[] PAPA_ODDS = {0.0, 0.67267686, 1.5146198, 0.27887323, 0.5460614, 2.313433, 0.96153843, 0.75686276,
            2.2562358, 0.20664589, 0.9607843, 1.9615384, 1.0836071, 0.30196398, 1.0716166, 0.6664044, 1.1432927, 0.8917492,
            2.2562358, 1.9478673, 2.1785367, 0.0};
        static final double[] BG_SCER = {0, 0.0550, 0.0126, 0.0586, 0.0655, 0.0441, 0.0498, 0.0217, 0.0655, 0.0735, 0.0950,
            0.0207, 0.0615, 0.0438, 0.0396, 0.0444, 0.0899, 0.0592, 0.0556, 0.0104, 0.0337, 0};
        static final double[] PRD_28 = {0, 0.04865, 0.00219, 0.01638, 0.00783, 0.02537, 0.07603, 0.0181, 0.02018, 0.01641,
            0.02639, 0.02975, 0.25885, 0.05126, 0.15178, 0.025, 0.10988, 0.03841, 0.01972, 0.00157, 0.05624, 0};

        static double[] normalized(double[] a) {
            double total = 0;
            for (double x : a) total = total + x;
            if (total < 1e-12) total = 1;
            double[] r = new double[a.length];
            for (int i = 0; i < a.length; i++) r[i] = a[i] / total;
            return r;
        }

        void derive(double[] bgCounts, double[] fgFreq) {
            for (int i = 0; i < LUT; i++) logLut[i] = Math.log(1.0 + Math.exp(-i / 100.0));
            for (int k = 1; k <= 20; k++) papaLod[k] = Math.log(PAPA_ODDS[k]);
            final double ninth = 1.0 / 9.0;
            for (int k = 0; k < NAA; k++) hydro2[k] = ninth * HYDRO[k] + 0.5;
            bgScer = normalized(BG_SCER);
            double[] f = (fgFreq != null ? fgFreq : PRD_28).clone();
            f[0] = 0;
            f[21] = 0;
            double[] fgn = normalized(f);
            double[] b = (bgCounts != null ? bgCounts : new double[NAA]).clone();
            b[0] = 0;
            b[21] = 0;
            bgInput = normalized(b);
            double[] mix = new double[NAA];
            for (int i = 0; i < NAA; i++) mix[i] = alpha * bgScer[i] + (1 - alpha) * bgInput[i];
            double[] combo = normalized(mix);
            final double eps = 0.00001;   // pseudo-frequency of X and *
            fgn[0] = eps;
            fgn[21] = eps;
            combo[0] = eps;
            combo[21] = eps;
            fg = normalized(fgn);
            bg = normalized(combo);
            for (int j = 1; j < 21; j++) llr[j] = Math.log(fg[j] / bg[j]);
            // two-state model: background <-> prion-like; emissions normalised once more, all in log space
            double[] ebg = normalized(bg), efg = normalized(fg);
            double[][] t = {{99.9 / 100, 0.1 / 100}, {2.0 / 100, 98.0 / 100}};
            double[] init = {0.9524, 0.0476};
            boolean freeEnd = true;
            double[] fprob = new double[2];
            for (int i = 0; i < 2; i++) {
                double rs = 0;
                for (int j = 0; j < 2; j++) {
                    lt[i][j] = Math.log(t[i][j]);
                    rs = rs + t[i][j];
                }
                li[i] = Math.log(init[i]);
                fprob[i] = Math.max(0.0, 1.0 - rs);
                if (fprob[i] > 0.0001) freeEnd = false;
            }
            for (int i = 0; i < 2; i++) lf[i] = Math.log(freeEnd ? 1.0 : fprob[i]);
            for (int j = 0; j < NAA; j++) {
                le[0][j] = Math.log(ebg[j]);
                le[1][j] = Math.log(efg[j]);
            }
        }
    }

    // ------------------------------------------------------------------------------------------ FASTA reader
    /** The jar's reader semantics: readLine line ends; sequence lines are not trimmed; an empty line ends the record and
     *  everything up to the next '>' line is skipped; a name found while skipping is trimmed, one found at the end of
     *  the previous record is not. */
    static final class Fasta implements AutoCloseable {
        private BufferedReader in;
        private boolean onDeck;
        String name;

        Fasta(String path) {
            try {
                in = new BufferedReader(new FileReader(path));
            } catch (IOException e) {
                System.out.println("# Couldn't open " + path);
            }
        }

        boolean hasMore() throws IOException {
            if (in == null) return false;
            if (onDeck) return true;
            String line;
            while ((line = in.readLine()) != null) {
                if (line.startsWith(">")) {
                    name = line.trim().substring(1);
                    return true;
                }
            }
            return false;
        }

        String next() throws IOException {
            StringBuilder sb = new StringBuilder();
            String line;
            onDeck = false;
            while ((line = in.readLine()) != null) {
                if (line.isEmpty()) return sb.toString();
                if (line.startsWith(">")) {
                    onDeck = true;
                    String seq = sb.toString();
                    pending = line.substring(1);
                    return seq;
                }
                sb.append(line);
            }
            return sb.toString();
        }

        private String pending;

        /** Call after next(): moves the name of the record found at the end of the previous one into place. */
        void advance() {
            if (onDeck) name = pending;
        }

        @Override
        public void close() throws IOException {
            if (in != null) in.close();
        }
    }

    static byte code(char c) {
        int k = AA.indexOf(Character.toUpperCase(c));
        return (byte) (k < 0 ? 0 : k);     // anything outside the alphabet scores as X
    }

    /** One scoring batch: residue codes of all proteins back to back (terminal '*' already stripped). */
    static final class Batch {
        final List<String> names = new ArrayList<>(), ids = new ArrayList<>();
        final List<Long> offsets = new ArrayList<>(List.of(0L));
        byte[] codes = new byte[1 << 20];
        int ncodes = 0;

        void add(String name, String id, String seq) {
            if (ncodes + seq.length() > codes.length) codes = java.util.Arrays.copyOf(codes, Math.max(2 * codes.length, ncodes + seq.length()));
            for (int i = 0; i < seq.length(); i++) codes[ncodes++] = code(seq.charAt(i));
            offsets.add((long) ncodes);
            names.add(name);
            ids.add(id);
        }

        void clear() {
            names.clear();
            ids.clear();
            offsets.clear();
            offsets.add(0L);
            ncodes = 0;
        }
    }

    // ------------------------------------------------------------------------------------------ background counts
    /** Residue counts of the proteins of a FASTA that have no X or * inside and no X at the end (on the UNSTRIPPED
     *  sequence); 64-bit counters, where the reference counts in int. */
    static double[] countBackground(String path) throws IOException {
        long[] cnt = new long[NAA];
        try (Fasta f = new Fasta(path)) {
            while (f.hasMore()) {
                String seq = f.next();
                f.advance();
                int m = seq.length();
                if (m == 0) continue;
                boolean valid = code(seq.charAt(m - 1)) != 0;
                for (int i = 1; valid && i + 1 < m; i++) {
                    byte c = code(seq.charAt(i));
                    valid = c != 0 && c != 21;
                }
                if (!valid) continue;
                for (int i = 0; i < m; i++) cnt[code(seq.charAt(i))]++;
            }
        }
        double[] out = new double[NAA];
        for (int i = 0; i < NAA; i++) out[i] = cnt[i];
        return out;
    }

    /** 22 lines "value [# name]". */
    static double[] readFrequencies(String path) {
        double[] v = new double[NAA];
        try (BufferedReader r = new BufferedReader(new FileReader(path))) {
            String line;
            for (int i = 0; i < NAA && (line = r.readLine()) != null; i++) {
                String[] tok = line.trim().split("\\s+");
                if (tok.length >= 1 && !tok[0].isEmpty()) v[i] = Double.parseDouble(tok[0]);
                if (tok.length >= 3 && tok[2].charAt(0) != AA.charAt(i))
                    System.out.println("# warning: " + path + " does not have expected name in line" + (i + 1));
            }
        } catch (IOException e) {
            System.out.println("# Couldn't open " + path);
        }
        return v;
    }

    static String vec(double[] v) {
        StringBuilder sb = new StringBuilder();
        for (int i = 0; i < NAA; i++) sb.append(AA.charAt(i)).append('=').append(String.format("%.5f", v[i])).append(';');
        return sb.toString();
    }

    /** Substring of a protein with the clamping rules of the reference's submatrix(int[], r1, r2). */
    static String sub(byte[] codes, long lo, int m, int r1, int r2) {
        if (m <= 0) return "";
        if (r1 < 0) r1 = 0;
        if (r2 < r1) r2 = r1;
        if (r1 >= m) r1 = m - 1;
        if (r2 >= m) r2 = m - 1;
        StringBuilder sb = new StringBuilder();
        for (int i = r1; i <= r2; i++) sb.append(AA.charAt(codes[(int) lo + i]));
        return sb.toString();
    }

    static double inf2nan(double x) {
        return Double.isInfinite(x) ? Double.NaN : x;
    }

    static final String SUMMARY_HEADER = "SEQid\tMW\tMWstart\tMWend\tMWlen\tLLR\tLLRstart\tLLRend\tLLRlen\tNLLR\tVITmaxrun\t"
        + "COREscore\tCOREstart\tCOREend\tCORElen\tPRDscore\tPRDstart\tPRDend\tPRDlen\tPROTlen\tHMMall\tHMMvit\tCOREaa\tSTARTaa\t"
        + "ENDaa\tPRDaa\tFInumaa\tFImeanhydro\tFImeancharge\tFImeancombo\tFImaxrun\tPAPAcombo\tPAPAprop\tPAPAfi\tPAPAllr\t"
        + "PAPAllr2\tPAPAcen\tPAPAaa";
    static final String RESIDUE_HEADER = "ORDER\tSEQid\tAANUM\tAA\tVIT\tMAP\tCHARGE\tHYDRO\tFI\tPLAAC\tPAPA\tFIx2\tPLAACx2\t"
        + "PAPAx2\tHMM.background\tHMM.PrD-like";

    // ------------------------------------------------------------------------------------------ the two tables
    static void printSummaryBatch(List<Cuda> cudas, boolean ranked, Batch b, Params p, PrintStream out) throws Throwable {
        if (b.names.isEmpty()) return;
        try (Arena a = Arena.ofConfined()) {
            Cuda cuda = cudas.get(0);
            MemorySegment s = cudas.size() > 1 ? Cuda.scoreMulti(cudas, b, a) : cuda.score(b, a);
            int[] rowOrder = ranked ? cuda.rank(s, b.names.size(), a) : null;    // (ranking is per batch: use one batch)
            for (int row = 0; row < b.names.size(); row++) {
                int i = ranked ? rowOrder[row] : row;
                long r = i * Cuda.SUMMARY_BYTES;
                int[] I = new int[14];
                for (int k = 0; k < 14; k++) I[k] = s.get(JAVA_INT, r + 4L * k);
                double[] D = new double[13];
                for (int k = 0; k < 13; k++) D[k] = s.get(JAVA_DOUBLE, r + 56 + 8L * k);
                int mwScore = I[0], mwStart = I[1], mwEnd = I[2], llrStart = I[3], llrEnd = I[4], maxRun = I[5];
                int coreStart = I[6], coreEnd = I[7], prdStart = I[8], prdEnd = I[9], n = I[10], fiNum = I[11], fiMax = I[12];
                int papaCen = I[13];
                if (n < 1) continue;
                long lo = b.offsets.get(i);
                double llr = inf2nan(D[0]);
                int llrLen = llrEnd - llrStart + 1, prdLen = prdEnd - prdStart + 1;
                out.format("%s\t%d\t%d\t%d\t%d\t%.3f\t%d\t%d\t%d\t%.3f\t%d\t%.3f\t%d\t%d\t%d\t%.3f\t%d\t%d\t%d\t%d\t%.3f\t%.3f",
                    b.names.get(i), mwScore, mwStart + 1, mwEnd + 1, mwEnd - mwStart + 1, llr, llrStart + 1, llrEnd + 1, llrLen,
                    llr / llrLen, maxRun, inf2nan(D[1]), coreStart + 1, coreEnd + 1, coreEnd - coreStart + 1, D[2], prdStart + 1,
                    prdEnd + 1, prdLen, n, D[3], D[4]);
                if (prdLen >= p.coreLength) {
                    out.print("\t" + sub(b.codes, lo, n, coreStart, coreEnd) + "\t" + sub(b.codes, lo, n, prdStart, prdStart + 14)
                        + "\t" + sub(b.codes, lo, n, prdEnd - 14, prdEnd) + "\t" + sub(b.codes, lo, n, prdStart, prdEnd));
                } else {
                    out.print("\t-\t-\t-\t-");
                }
                out.format("\t%d\t%.3f\t%.3f\t%.3f\t%d\t%.3f\t%.3f\t%.3f\t%.3f\t%.3f\t%d\t%s\n", fiNum, D[5], D[6], D[7], fiMax,
                    inf2nan(D[8]), D[9], D[10], D[11], D[12], papaCen + 1,
                    sub(b.codes, lo, n, papaCen - p.ww2 / 2, papaCen + p.ww2 / 2));
            }
        }
        b.clear();
    }

    static void printResidueBatch(Cuda cuda, Batch b, PrintStream out) throws Throwable {
        if (b.names.isEmpty()) return;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment[] r = cuda.scoreResidues(b, a);
            for (int i = 0; i < b.names.size(); i++) {
                long lo = b.offsets.get(i), hi = b.offsets.get(i + 1);
                for (long t = lo; t < hi; t++) {
                    out.print(b.ids.get(i) + "\t" + b.names.get(i) + "\t" + (t - lo + 1) + "\t" + AA.charAt(b.codes[(int) t]) + "\t"
                        + r[0].get(JAVA_BYTE, t) + "\t" + r[1].get(JAVA_BYTE, t) + "\t");
                    out.format("%.4f\t%.4f\t%.8f\t%.4f\t%.8f\t%.8f\t%.4f\t%.8f", r[2].getAtIndex(JAVA_DOUBLE, t),
                        r[3].getAtIndex(JAVA_DOUBLE, t), r[4].getAtIndex(JAVA_DOUBLE, t), r[5].getAtIndex(JAVA_DOUBLE, t),
                        r[6].getAtIndex(JAVA_DOUBLE, t), r[7].getAtIndex(JAVA_DOUBLE, t), r[8].getAtIndex(JAVA_DOUBLE, t),
                        r[9].getAtIndex(JAVA_DOUBLE, t));
                    out.format("\t%.4f\t%.4f\n", r[10].getAtIndex(JAVA_DOUBLE, t), r[11].getAtIndex(JAVA_DOUBLE, t));
                }
                out.println("########################################################");
            }
        }
        b.clear();
    }

    // ------------------------------------------------------------------------------------------ main
    public static void main(String[] argv) throws Throwable {
        String input = "", bgFasta = "", bgFreqFile = "", fgFreqFile = "", plotList = "";
        Params p = new Params();
        boolean printDocs = false, printParams = true, compatF = false, ranked = false;
        int device = 0, gpus = 1;
        long batchResidues = 256L << 20;
        List<String> args = new ArrayList<>();
        for (int i = 0; i < argv.length; i++) {          // options of this host start with "--"
            switch (argv[i]) {
                case "--device" -> device = Integer.parseInt(argv[++i]);
                case "--gpus" -> gpus = Integer.parseInt(argv[++i]);       // one ctx + host thread per GPU (summary table)
                case "--rank" -> ranked = true;                            // rows in the web front end's order
                case "--batch-mb" -> batchResidues = Long.parseLong(argv[++i]) << 20;
                case "--compat-F" -> compatF = true;
                default -> args.add(argv[i]);
            }
        }
        // the jar consumes options pairwise and looks at a trailing token only if it is -d or -s
        int i = 0, n = args.size();
        while (i + 1 < n || (i < n && (args.get(i).equals("-d") || args.get(i).equals("-s")))) {
            String a = args.get(i);
            switch (a) {
                case "-i" -> input = args.get(++i);
                case "-b" -> bgFasta = args.get(++i);
                case "-B" -> bgFreqFile = args.get(++i);
                case "-F" -> fgFreqFile = args.get(++i);
                case "-c" -> p.coreLength = Integer.parseInt(args.get(++i));
                case "-w" -> p.ww1 = Integer.parseInt(args.get(++i));
                case "-W" -> p.ww2 = Integer.parseInt(args.get(++i));
                case "-a" -> p.alpha = Double.parseDouble(args.get(++i));
                case "-m", "-h" -> ++i;                    // hmm type is unused by the jar; the GraphViz export is not part of this host
                case "-p" -> plotList = args.get(++i);
                case "-d" -> printDocs = true;
                case "-s" -> printParams = false;
                default -> System.out.println("# skipping unknown option " + a);
            }
            i++;
        }
        p.ww3 = p.ww2;

        double[] bgCounts = new double[NAA];
        if (!bgFreqFile.isEmpty()) bgCounts = readFrequencies(bgFreqFile);
        else if (!bgFasta.isEmpty()) bgCounts = countBackground(bgFasta);
        else if (!input.isEmpty()) bgCounts = countBackground(input);
        double[] fgFreq = null;
        if (!fgFreqFile.isEmpty()) fgFreq = readFrequencies(compatF ? bgFreqFile : fgFreqFile);   // the jar reads the -B file here
        if ((!bgFasta.isEmpty() || !bgFreqFile.isEmpty()) && input.isEmpty()) {
            for (int k = 0; k < NAA; k++) System.out.println(String.format("%.6f", bgCounts[k]) + " # " + AA.charAt(k));
            return;
        }
        if (input.isEmpty()) {
            System.out.println("USAGE: java Plaac -i input.fa [-c core_length] [-B bg_freqs.txt | -b background.fa] [-a alpha] "
                + "[-F fg_freqs.txt] [-w window] [-W Window] [-d] [-s] [-p list.txt|all] > output.txt");
            return;
        }
        if (p.alpha > 1 || p.alpha < 0) {
            System.out.println("# warning: invalid alpha; using alpha = 1.0");
            p.alpha = 1.0;
        }
        p.derive(bgCounts, fgFreq);
        PrintStream out = System.out;
        if (printParams) {
            out.println("############################ parameters at run-time ####################################");
            out.println("## alpha=" + p.alpha + "; corelength=" + p.coreLength + "; ww1=" + p.ww1 + "; ww2=" + p.ww2 + "; ww3=" + p.ww3
                + "; adjustprolines=" + p.adjustProlines + ";");
            out.println("## fg_used: {" + vec(p.fg) + "}");
            out.println("## bg_scer: {" + vec(p.bgScer) + "}");
            out.println("## bg_input: {" + vec(p.bgInput) + "}");
            out.println("## bg_used: {" + vec(p.bg) + "}");
            out.println("## plaac_llr: {" + vec(p.llr) + "}");
            out.println("## papa_lods: {" + vec(p.papaLod) + "}");
            out.println("#######################################################################################");
        }
        if (Cuda.deviceCount() <= 0) {
            System.err.println("Plaac: no CUDA device (" + Cuda.lastError(MemorySegment.NULL) + "); there is no CPU scoring path");
            System.exit(2);
        }
        List<Cuda> cudas = new ArrayList<>();
        gpus = Math.max(1, Math.min(gpus, Cuda.deviceCount() - device));
        try (Cuda cuda = new Cuda(device, p); Fasta f = new Fasta(input)) {
            cudas.add(cuda);
            for (int g = 1; g < gpus; g++) cudas.add(new Cuda(device + g, p));
            if (ranked) batchResidues = Long.MAX_VALUE;   // one batch, so that the order is that of the whole table
            Batch b = new Batch();
            if (plotList.isEmpty()) {
                if (printDocs) ColumnDocs.print(out);
                out.println(SUMMARY_HEADER);
                while (f.hasMore()) {
                    String name = f.name;
                    String seq = f.next();
                    f.advance();
                    if (seq.endsWith("*")) seq = seq.substring(0, seq.length() - 1);
                    if (seq.isEmpty()) continue;
                    b.add(name, "", seq);
                    if (b.ncodes >= batchResidues || (!ranked && b.names.size() >= (4 << 20))) printSummaryBatch(cudas, ranked, b, p, out);
                }
                printSummaryBatch(cudas, ranked, b, p, out);
            } else {
                boolean all = plotList.equals("all");
                Map<String, String> synonym = new HashMap<>(), order = new HashMap<>();
                if (!all) {
                    try (BufferedReader r = new BufferedReader(new FileReader(plotList))) {
                        String line;
                        int k = 1;
                        while ((line = r.readLine()) != null) {
                            String[] tok = line.split("\t");
                            synonym.put(tok[0], tok.length > 1 ? tok[1] : tok[0]);
                            order.put(tok[0], Integer.toString(k++));
                        }
                    } catch (IOException e) {
                        out.println("# Couldn't open " + plotList);
                    }
                }
                out.println(RESIDUE_HEADER);
                int k = 0;
                while (f.hasMore()) {
                    String name = f.name;
                    String seq = f.next();
                    f.advance();
                    k++;
                    if (!all && !synonym.containsKey(name)) continue;
                    if (seq.endsWith("*")) seq = seq.substring(0, seq.length() - 1);
                    if (seq.isEmpty()) continue;
                    b.add(all ? name : synonym.get(name), all ? Integer.toString(k) : order.get(name), seq);
                    if (b.ncodes >= (32 << 20)) printResidueBatch(cuda, b, out);
                }
                printResidueBatch(cuda, b, out);
            }
        } finally {
            for (int g = 1; g < cudas.size(); g++) cudas.get(g).close();
        }
    }

    /** Text of the "-d" block: one line per output column, "## name: description" (38 columns). */
    static final class ColumnDocs {
        static void print(PrintStream out) {
            out.println("############################ Description of output columns ############################");
            // The 38 description lines are data shared with the C++ host (kColumnDocs in plaac_cli.cpp, checked there
            // against the reference's generated web/views/_plaac_headers.haml); they are read from the resource that
            // ships next to this class so that the two hosts cannot drift apart.
            try (BufferedReader r = new BufferedReader(new java.io.InputStreamReader(
                    Plaac.class.getResourceAsStream("column_docs.txt"), java.nio.charset.StandardCharsets.UTF_8))) {
                String line;
                while ((line = r.readLine()) != null) out.println("## " + line);
            } catch (IOException | NullPointerException e) {
                out.println("## (column_docs.txt not found next to Plaac.class)");
            }
            out.println("#######################################################################################");
        }
    }
}
