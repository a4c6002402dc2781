# One HGX B200 search node + N CPU crawler VMs on AWS.  `terraform apply -var admin_cidr=203.0.113.0/24`
terraform {
  required_version = ">= 1.6"
  required_providers { aws = { source = "hashicorp/aws", version = "~> 5.0" } }
}

variable "region"        { default = "us-east-1" }
variable "admin_cidr"    { description = "CIDR allowed to reach SSH / the admin API" }
variable "crawler_count" { default = 2 }
variable "image"         { default = "ghcr.io/example/infomesh-b200:latest" }
variable "api_key"       { sensitive = true }
variable "search_instance_type"  { default = "p6-b200.48xlarge" }   # 8x B200, NVSwitch
variable "crawler_instance_type" { default = "c7i.large" }

provider "aws" { region = var.region }

data "aws_ami" "dl" {
  most_recent = true
  owners      = ["amazon"]
  filter {
    name   = "name"
    values = ["Deep Learning Base OSS Nvidia Driver GPU AMI (Ubuntu 24.04)*"]
  }
}

resource "aws_vpc" "mesh" { cidr_block = "10.42.0.0/16" }
resource "aws_internet_gateway" "gw" { vpc_id = aws_vpc.mesh.id }
resource "aws_subnet" "private" {
  vpc_id     = aws_vpc.mesh.id
  cidr_block = "10.42.1.0/24"
}
resource "aws_subnet" "dmz" {
  vpc_id                  = aws_vpc.mesh.id
  cidr_block              = "10.42.2.0/24"
  map_public_ip_on_launch = true
}
resource "aws_route_table" "public" {
  vpc_id = aws_vpc.mesh.id
  route {
    cidr_block = "0.0.0.0/0"
    gateway_id = aws_internet_gateway.gw.id
  }
}
resource "aws_route_table_association" "dmz" {
  subnet_id      = aws_subnet.dmz.id
  route_table_id = aws_route_table.public.id
}

resource "aws_security_group" "search" {
  vpc_id = aws_vpc.mesh.id
  ingress {                                     # crawlers submit pages and gossip over the private subnet only
    from_port   = 0
    to_port     = 65535
    protocol    = "tcp"
    cidr_blocks = [aws_subnet.dmz.cidr_block]
  }
  ingress {
    from_port   = 8080
    to_port     = 8081
    protocol    = "tcp"
    cidr_blocks = [var.admin_cidr]
  }
  egress {
    from_port   = 0
    to_port     = 0
    protocol    = "-1"
    cidr_blocks = ["0.0.0.0/0"]
  }
}
resource "aws_security_group" "crawler" {
  vpc_id = aws_vpc.mesh.id
  ingress {
    from_port   = 4001
    to_port     = 4001
    protocol    = "tcp"
    cidr_blocks = ["0.0.0.0/0"]
  }
  ingress {
    from_port   = 22
    to_port     = 22
    protocol    = "tcp"
    cidr_blocks = [var.admin_cidr]
  }
  egress {
    from_port   = 0
    to_port     = 0
    protocol    = "-1"
    cidr_blocks = ["0.0.0.0/0"]
  }
}

locals {
  run = "docker run -d --restart unless-stopped --name infomesh -v /data:/data -e INFOMESH_API_KEY=${var.api_key}"
}

resource "aws_instance" "search" {
  ami                    = data.aws_ami.dl.id
  instance_type          = var.search_instance_type
  subnet_id              = aws_subnet.private.id
  vpc_security_group_ids = [aws_security_group.search.id]
  root_block_device { volume_size = 1000 }
  user_data = <<-EOT
    #!/bin/bash
    mkdir -p /data && chown 1000:1000 /data
    ${local.run} --gpus all --shm-size 16g --ulimit memlock=-1 -p 4001:4001 -p 8080:8080 -p 8081:8081 \
      -e INFOMESH_NODE_ROLE=search -e INFOMESH_GPU_ENABLED=true ${var.image}
  EOT
  tags = { Name = "infomesh-search" }
}

resource "aws_instance" "crawler" {
  count                  = var.crawler_count
  ami                    = data.aws_ami.dl.id
  instance_type          = var.crawler_instance_type
  subnet_id              = aws_subnet.dmz.id
  vpc_security_group_ids = [aws_security_group.crawler.id]
  user_data = <<-EOT
    #!/bin/bash
    mkdir -p /data && chown 1000:1000 /data
    ${local.run} -p 4001:4001 -e INFOMESH_NODE_ROLE=crawler -e INFOMESH_GPU_ENABLED=false \
      -e INFOMESH_NETWORK_INDEX_SUBMIT_PEERS=http://${aws_instance.search.private_ip}:8080 ${var.image}
  EOT
  tags = { Name = "infomesh-crawler-${count.index}" }
}

output "search_private_ip" { value = aws_instance.search.private_ip }
output "crawler_public_ips" { value = aws_instance.crawler[*].public_ip }
