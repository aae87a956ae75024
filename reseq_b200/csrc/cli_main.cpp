// reseq-b200: command line shell over libreseq_b200.so that keeps the reference's `reseq illuminaPE` /
// `reseq seqToIllumina` simulation options (reference reseq/main.cpp:694-1137; README.md:133-231) for the hot path.
// Profile creation (-b/--bamIn) and IPF fitting are outside this path: those options are rejected with a message instead of
// being silently ignored. -V/--vcfSim files are read and checked like the reference does; simulating with them is refused by
// the engine until its variant-aware kernels exist.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include "../../include/reseq_b200.h"

namespace {
int g_verbosity = 4;
void info(const std::string &m){ if(g_verbosity >= 3){ fprintf(stderr, ">>> info:  %s\n", m.c_str()); } }
int err(const std::string &m){ if(g_verbosity >= 1){ fprintf(stderr, "!!! error: %s\n", m.c_str()); } return 1; }

struct Args {
	std::map<std::string, std::string> kv;
	std::map<std::string, bool> flags;
	bool has(const std::string &k) const { return kv.count(k) != 0; }
	std::string get(const std::string &k, const std::string &d = "") const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
};

// long name -> takes a value?; short aliases as in main.cpp:710-754, 1009-1021
const std::map<std::string, bool> kOpts = {
	{"threads", true}, {"verbosity", true}, {"statsIn", true}, {"probabilitiesIn", true}, {"ipfIterations", true}, {"ipfPrecision", true},
	{"firstReadsOut", true}, {"secondReadsOut", true}, {"coverage", true}, {"numReads", true}, {"seed", true}, {"refBias", true},
	{"recordBaseIdentifier", true}, {"refSim", true}, {"refIn", true}, {"errorMutliplier", true}, {"noInDelErrors", false},
	{"noSubstitutionErrors", false}, {"input", true}, {"output", true}, {"gpus", true}, {"help", false},
	// recognised but unsupported on this path
	{"bamIn", true}, {"vcfSim", true}, {"methylation", true}, {"readSysError", true}, {"writeSysError", true}, {"refBiasFile", true},
	{"statsOnly", false}, {"stopAfterEstimation", false}, {"vcfIn", true}, {"adapterFile", true}, {"adapterMatrix", true},
};
const std::map<char, std::string> kShort = {
	{'j', "threads"}, {'s', "statsIn"}, {'p', "probabilitiesIn"}, {'1', "firstReadsOut"}, {'2', "secondReadsOut"}, {'c', "coverage"},
	{'R', "refSim"}, {'r', "refIn"}, {'i', "input"}, {'o', "output"}, {'b', "bamIn"}, {'V', "vcfSim"}, {'v', "vcfIn"}, {'h', "help"},
};

bool parse(int argc, char **argv, int first, Args &a){
	for(int i = first; i < argc; ++i){
		std::string t = argv[i], name;
		if(t.rfind("--", 0) == 0){ name = t.substr(2); }
		else if(t.size() == 2 && t[0] == '-' && kShort.count(t[1])){ name = kShort.at(t[1]); }
		else{ err("unrecognised option '" + t + "'"); return false; }
		auto it = kOpts.find(name);
		if(it == kOpts.end()){ err("unrecognised option '" + t + "'"); return false; }
		if(it->second){
			if(i + 1 >= argc){ err("the required argument for option '" + t + "' is missing"); return false; }
			a.kv[name] = argv[++i];
		}
		else{ a.flags[name] = true; }
	}
	return true;
}

void usage(){
	fprintf(stderr,
	        "Usage:  reseq-b200 illuminaPE -s <stats.reseq> -R <ref.fa> -1 <R1.fq> -2 <R2.fq> [options]\n"
	        "        reseq-b200 seqToIllumina -i <frags.fa> -o <out.fq> -s <stats.reseq> [options]\n"
	        "Simulation options (as in reseq): -c/--coverage, --numReads, --seed, --refBias keep|no, --recordBaseIdentifier,\n"
	        "  -p/--probabilitiesIn, --errorMutliplier, --noInDelErrors, --noSubstitutionErrors, -j (ignored), --verbosity\n"
	        "Extra: --gpus N   split the SimBlocks over N devices (output is identical to N=1)\n");
}

int reject_unsupported(const Args &a){
	for(const char *k : {"bamIn", "vcfIn", "adapterFile", "adapterMatrix"}){
		if(a.has(k)){ return err(std::string("option --") + k + " belongs to a part of ReSeq this engine does not replace (stats creation); run it with the reference implementation"); }
	}
	for(const char *k : {"statsOnly", "stopAfterEstimation"}){ if(a.flags.count(k)){ return err(std::string("option --") + k + " is a stats/IPF step; use the reference implementation"); } }
	if(a.has("ipfIterations") && a.get("ipfIterations") != "0"){ return err("this engine does not fit probabilities: pass a converged X.reseq.ipf (and --ipfIterations 0)"); }
	return 0;
}

rsq_profile *load_profile(const Args &a){
	const std::string stats = a.get("statsIn");
	if(stats.empty()){ err("statsIn option is mandatory (stats creation from bam is not part of this engine)."); return nullptr; }
	rsq_profile *p = nullptr;
	if(stats.size() > 5 && stats.substr(stats.size() - 5) == ".flat"){ p = rsq_profile_load_flat(stats.c_str()); }
	else{
		const std::string ipf = a.get("probabilitiesIn", stats + ".ipf");
		info("Reading real data statistics from " + stats);
		info("Loading probability estimates from '" + ipf + "'");
		p = rsq_profile_load(stats.c_str(), ipf.c_str());
	}
	if(!p){ err(rsq_last_error()); return nullptr; }
	if(a.flags.count("noInDelErrors")){ info("Removing InDel errors"); if(rsq_profile_remove_indel_errors(p)){ err(rsq_last_error()); } }
	if(a.flags.count("noSubstitutionErrors")){ info("Removing substitution errors"); rsq_profile_remove_substitution_errors(p); }
	if(a.has("errorMutliplier")){
		const double m = atof(a.get("errorMutliplier").c_str());
		if(m == 0.0){ rsq_profile_remove_substitution_errors(p); }
		else if(m != 1.0){ if(rsq_profile_change_error_rate(p, m)){ err(rsq_last_error()); rsq_profile_free(p); return nullptr; } }
	}
	return p;
}

uint64_t pick_seed(const Args &a){
	if(a.has("seed")){ return strtoull(a.get("seed").c_str(), nullptr, 10); }
	std::random_device rd;
	const uint64_t s = rd();
	info("Randomly generated seed is " + std::to_string(s));
	return s;
}

int run_illumina_pe(const Args &a){
	if(int rc = reject_unsupported(a)){ return rc; }
	const std::string ref_path = a.get("refSim", a.get("refIn"));
	if(ref_path.empty()){ return err("refIn or refSim option mandatory."); }
	const std::string out1 = a.get("firstReadsOut", "reseq-R1.fq"), out2 = a.get("secondReadsOut", "reseq-R2.fq");
	info("Reading reference from " + ref_path);
	rsq_reference *ref = rsq_reference_load_fasta(ref_path.c_str());
	if(!ref){ return err(rsq_last_error()); }
	// -V before --methylation like Simulator::Simulate opens them (Simulator.cpp:2747-2772): the methylation file may carry one column per allele
	if(a.has("vcfSim") && rsq_reference_load_variants(ref, a.get("vcfSim").c_str())){ rsq_reference_free(ref); return err(rsq_last_error()); }
	if(a.has("methylation") && rsq_reference_load_methylation(ref, a.get("methylation").c_str())){ rsq_reference_free(ref); return err(rsq_last_error()); }
	rsq_profile *prof = load_profile(a);
	if(!prof){ rsq_reference_free(ref); return 1; }
	rsq_sim_options opt{};
	opt.seed = pick_seed(a);
	opt.coverage = atof(a.get("coverage", "0").c_str());
	opt.num_read_pairs = strtoull(a.get("numReads", "0").c_str(), nullptr, 10);
	const std::string bias_file = a.get("refBiasFile");
	const std::string bias = a.get("refBias", a.has("refBiasFile") ? "file" : (a.has("refSim") ? "no" : "keep"));   // main.cpp:863-908
	if(bias == "keep"){ opt.ref_bias_model = 0; } else if(bias == "no"){ opt.ref_bias_model = 1; } else if(bias == "draw"){ opt.ref_bias_model = 2; }
	else if(bias == "file"){ opt.ref_bias_model = 3; }
	else{ rsq_profile_free(prof); rsq_reference_free(ref); return err("Unknown option for refBias: " + bias); }
	if(opt.ref_bias_model == 3 && bias_file.empty()){ rsq_profile_free(prof); rsq_reference_free(ref); return err("refBiasFile option mandatory if for refBias option file was chosen"); }
	if(opt.ref_bias_model != 3 && !bias_file.empty()){ rsq_profile_free(prof); rsq_reference_free(ref); return err("refBiasFile option only allowed if for refBias option file was chosen"); }
	opt.ref_bias_file = bias_file.empty() ? nullptr : bias_file.c_str();
	const std::string base_id = a.get("recordBaseIdentifier", "ReseqRead");
	opt.record_base_identifier = base_id.c_str();
	int gpus = atoi(a.get("gpus", "1").c_str());
	if(gpus < 1){ gpus = 1; }
	if(gpus > rsq_device_count()){ rsq_profile_free(prof); rsq_reference_free(ref); return err("requested " + std::to_string(gpus) + " GPUs but only " + std::to_string(rsq_device_count()) + " CUDA devices are usable (there is no CPU path)"); }
	if(a.has("readSysError") && a.has("writeSysError")){ rsq_profile_free(prof); rsq_reference_free(ref); return err("writeSysError and readSysError option are mutually exclusive. Specify the one or the other."); }
	std::string sys_error_file = a.get("readSysError");
	if(a.has("writeSysError")){
		// main.cpp:379-391: create the profile with this seed, then simulate from the file
		sys_error_file = a.get("writeSysError");
		rsq_engine *w = rsq_engine_create(prof, 0);
		if(!w || rsq_create_systematic_error_profile(w, ref, opt.seed, sys_error_file.c_str())){
			if(w){ rsq_engine_destroy(w); }
			rsq_profile_free(prof); rsq_reference_free(ref);
			return err(rsq_last_error());
		}
		rsq_engine_destroy(w);
	}
	opt.sys_error_file = sys_error_file.empty() ? nullptr : sys_error_file.c_str();
	info("Storing simulated data in " + out1 + " and " + out2);
	for(const std::string &o : {out1, out2}){ FILE *f = fopen(o.c_str(), "wb"); if(!f){ return err("Could not open '" + o + "' for writing."); } fclose(f); }

	// the drop-in call streams batch after batch into the two files (runs larger than HBM or host memory work); with several GPUs every engine
	// streams its shard and the shard files are appended in order (rsq_simulate_multi)
	const auto t1 = std::chrono::steady_clock::now();
	rsq_sim_report rep;
	const int rc1 = gpus == 1 ? rsq_simulate(prof, ref, &opt, 0, out1.c_str(), out2.c_str(), &rep)
	                          : rsq_simulate_multi(prof, ref, &opt, gpus, nullptr, out1.c_str(), out2.c_str(), &rep);
	const std::string msg = rc1 ? rsq_last_error() : "";
	rsq_profile_free(prof); rsq_reference_free(ref);
	if(rc1){ err(msg); err("An error occurred in the process: Terminating simulation"); return 1; }
	const double secs1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
	info("Generated " + std::to_string(rep.pairs) + " read pairs (aim " + std::to_string(rep.total_pairs_aim) + ") in " + std::to_string(secs1) + " s on " + std::to_string(gpus) + " GPU(s).");
	info("Simulation finished succesfully");
	return 0;
}

int run_seq_to_illumina(const Args &a){
	if(int rc = reject_unsupported(a)){ return rc; }
	if(!a.has("input") || !a.has("output")){ return err("this build needs -i <input.fa> and -o <output.fq> (stdin/stdout streaming is not supported)"); }
	rsq_profile *prof = load_profile(a);
	if(!prof){ return 1; }
	rsq_engine *e = rsq_engine_create(prof, 0);
	if(!e){ rsq_profile_free(prof); return err(rsq_last_error()); }
	rsq_sim_report rep;
	int rc = rsq_apply_error_model(e, a.get("input").c_str(), a.get("output").c_str(), pick_seed(a), &rep);
	if(rc){ err(rsq_last_error()); err("An error occurred in the process: Terminating simulation"); }
	else{ info("Generated " + std::to_string(rep.pairs) + " reads."); info("Simulation finished succesfully"); }
	rsq_engine_destroy(e); rsq_profile_free(prof);
	return rc ? 1 : 0;
}
}  // namespace

int main(int argc, char **argv){
	if(argc < 2){ usage(); return 1; }
	const std::string cmd = argv[1];
	Args a;
	if(!parse(argc, argv, 2, a)){ usage(); return 1; }
	if(a.has("verbosity")){ g_verbosity = atoi(a.get("verbosity").c_str()); }
	if(a.flags.count("help")){ usage(); return 0; }
	if(cmd == "illuminaPE"){ return run_illumina_pe(a); }
	if(cmd == "seqToIllumina"){ return run_seq_to_illumina(a); }
	err("command '" + cmd + "' is not part of this engine (illuminaPE, seqToIllumina)");
	usage();
	return 1;
}
