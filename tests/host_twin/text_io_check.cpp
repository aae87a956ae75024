// TEST HARNESS for reseq_b200/csrc/text_io.hpp (host text I/O either side of the simulation path).
//   text_io_check compress <in> <out[.gz]> <piece bytes>   - writes <in> through TextSink in pieces (piece 0 = one write)
//   text_io_check cat <in[.gz]>                            - streams a (possibly gzip) file through TextInput to stdout
//   text_io_check empty <out.gz>                           - opens and closes a sink without writing
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include "../../reseq_b200/csrc/text_io.hpp"

int main(int argc, char **argv){
	try{
		const std::string mode = argc > 1 ? argv[1] : "";
		if(mode == "compress" && argc == 5){
			std::ifstream f(argv[2], std::ios::binary);
			const std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
			rsq::TextSink sink;
			if(!sink.open(argv[3])){ fprintf(stderr, "cannot open %s\n", argv[3]); return 1; }
			size_t piece = static_cast<size_t>(atoll(argv[4]));
			if(!piece){ piece = text.size() ? text.size() : 1; }
			for(size_t off = 0; off < text.size(); off += piece){
				if(!sink.write(text.data() + off, std::min(piece, text.size() - off))){ fprintf(stderr, "write failed\n"); return 1; }
			}
			if(!sink.close()){ fprintf(stderr, "close failed\n"); return 1; }
			printf("text=%llu file=%llu compressed=%d\n", (unsigned long long)sink.text_bytes(), (unsigned long long)sink.file_bytes(), sink.compressed() ? 1 : 0);
			return 0;
		}
		if(mode == "cat" && argc == 3){
			rsq::TextInput in(argv[2]);
			if(!in.is_open()){ fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
			std::string line;
			while(std::getline(in.stream(), line)){ std::cout << line << '\n'; }
			if(in.corrupt()){ fprintf(stderr, "corrupt gzip stream\n"); return 2; }
			return 0;
		}
		if(mode == "empty" && argc == 3){
			rsq::TextSink sink;
			return sink.open(argv[2]) && sink.close() ? 0 : 1;
		}
	}
	catch(const std::exception &ex){ fprintf(stderr, "%s\n", ex.what()); return 3; }
	fprintf(stderr, "usage: text_io_check compress|cat|empty ...\n");
	return 64;
}
